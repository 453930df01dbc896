"""End-to-end parity of the CUDA path (product ResNet + FPN modules -> C ABI plan -> sm_100a
kernels) against the fp32 CPU oracle and the reference-generated golden vectors.

Gate (BASELINE.json north_star): every FPN level within relative L2 <= 1e-2 with bf16 I/O.
C2..C5 are checked to the same bound for diagnosis.
"""
import pytest
import torch

from oracle import resnet_fpn_oracle as orc
from tests import helpers

pytestmark = pytest.mark.gpu

BF16_GATE = 1e-2
FP32_GATE = 1e-4   # fp32 I/O: split-precision kernels (north_star tolerance)


def _run_product(bb, neck, x, dev):
    bb = bb.to(dev)
    neck = neck.to(dev)
    bb.eval()
    neck.eval()
    with torch.no_grad():
        feats = bb(x.to(dev))
        outs = neck(feats)
    torch.cuda.synchronize()
    return feats, outs


def _check_levels(got, want, names, gate=BF16_GATE):
    errs = {}
    for n, a, b in zip(names, got, want):
        assert tuple(a.shape) == tuple(b.shape), (n, a.shape, b.shape)
        errs[n] = orc.rel_l2(a.float(), b)
    bad = {k: v for k, v in errs.items() if not v <= gate}
    assert not bad, "rel-L2 over gate: %s (all: %s)" % (bad, errs)
    return errs


@pytest.mark.parametrize("name", helpers.GOLDEN_CASES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_golden_vectors(cuda_device, name, dtype):
    meta, arrays = helpers.load_golden(name)
    bb, neck = helpers.build_product_pair(meta["depth"], seed=meta["seed"], bnstats=bool(meta["bnstats"]))
    assert helpers.state_hash(bb.state_dict()) == meta["bb_hash"]
    feats, outs = _run_product(bb, neck, arrays["x"].to(dtype), cuda_device)
    assert all(t.dtype == dtype for t in feats + outs)
    gate = FP32_GATE if dtype == torch.float32 else BF16_GATE
    _check_levels(feats, [arrays["C%d" % i] for i in range(2, 6)], ["C2", "C3", "C4", "C5"], gate)
    e = _check_levels(outs, [arrays["P%d" % i] for i in range(2, 7)], ["P2", "P3", "P4", "P5", "P6"], gate)
    print("golden %s %s" % (name, dtype), e)


@pytest.mark.parametrize("depth,shape,bnstats", [
    (18, (2, 3, 128, 160), False),
    (34, (1, 3, 96, 128), True),
    (50, (2, 3, 128, 160), False),
    (50, (1, 3, 224, 320), True),
    (101, (1, 3, 128, 128), False),
])
def test_against_cpu_oracle(cuda_device, depth, shape, bnstats):
    torch.set_num_threads(max(1, torch.get_num_threads()))
    bb, neck = helpers.build_product_pair(depth, seed=11, bnstats=bnstats)
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(5))
    xb = x.to(torch.bfloat16)
    want_f, want_p = orc.resnet_fpn_forward(bsd, nsd, xb.float(), depth)
    feats, outs = _run_product(bb, neck, xb, cuda_device)
    e1 = _check_levels(feats, want_f, ["C2", "C3", "C4", "C5"])
    e2 = _check_levels(outs, want_p, ["P2", "P3", "P4", "P5", "P6"])
    print("rel-L2", depth, shape, e1, e2)


def test_full_size_r50_fpn_single_image(cuda_device):
    """BASELINE config geometry: one 800x1333 image zero-padded to 800x1344."""
    bb, neck = helpers.build_product_pair(50, seed=0)
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    g = torch.Generator().manual_seed(0)
    x = torch.zeros(1, 3, 800, 1344)
    x[:, :, :, :1333] = torch.randn(1, 3, 800, 1333, generator=g)
    xb = x.to(torch.bfloat16)
    want_f, want_p = orc.resnet_fpn_forward(bsd, nsd, xb.float(), 50)
    feats, outs = _run_product(bb, neck, xb, cuda_device)
    assert [tuple(t.shape) for t in outs] == [(1, 256, 200, 336), (1, 256, 100, 168), (1, 256, 50, 84),
                                              (1, 256, 25, 42), (1, 256, 13, 21)]
    e1 = _check_levels(feats, want_f, ["C2", "C3", "C4", "C5"])
    e2 = _check_levels(outs, want_p, ["P2", "P3", "P4", "P5", "P6"])
    print("rel-L2 full size", e1, e2)


def test_full_batch16_equals_single_image(cuda_device):
    """BASELINE config 2 at its full size (batch 16 @800x1344): every image of a batch of 16 copies reproduces the
    single-image result -- which `test_full_size_r50_fpn_single_image` checks against the oracle -- bit for bit.
    Size-independent property covering the whole-batch tile schedules (CTA pairs, fused stem + max-pool ranges that
    span images, 8 400-tile persistent grids)."""
    bb, neck = helpers.build_product_pair(50, seed=0)
    g = torch.Generator().manual_seed(0)
    x1 = torch.zeros(1, 3, 800, 1344)
    x1[:, :, :, :1333] = torch.randn(1, 3, 800, 1333, generator=g)
    x1 = x1.to(torch.bfloat16)
    f1, p1 = _run_product(bb, neck, x1, cuda_device)
    f16, p16 = _run_product(bb, neck, x1.expand(16, -1, -1, -1).contiguous(), cuda_device)
    for a, b in zip(f1 + p1, f16 + p16):
        assert b.shape[0] == 16
        for i in (0, 7, 15):
            assert torch.equal(a[0], b[i]), "image %d of the batch differs from the single-image run" % i
        assert torch.equal(b, b[0:1].expand_as(b))


def test_batch_independence_and_determinism(cuda_device):
    """Images are independent units (eval BN): batch of 3 == three batches of 1, bit for bit, and a
    second run reproduces the first exactly (size-independent property)."""
    bb, neck = helpers.build_product_pair(50, seed=2)
    x = torch.randn(3, 3, 96, 128, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16)
    f3, p3 = _run_product(bb, neck, x, cuda_device)
    f3b, p3b = _run_product(bb, neck, x, cuda_device)
    for a, b in zip(f3 + p3, f3b + p3b):
        assert torch.equal(a, b)
    for i in range(3):
        f1, p1 = _run_product(bb, neck, x[i:i + 1], cuda_device)
        for a, b in zip(f3 + p3, f1 + p1):
            assert torch.equal(a[i:i + 1], b)


def test_odd_size_backbone_and_fpn_mismatch_error(cuda_device):
    """Raw 1333-derived odd sizes run through the backbone; the FPN raises RuntimeError like the
    reference (fpn.py:100-101) instead of cropping."""
    bb, neck = helpers.build_product_pair(18, seed=0)
    bsd = helpers.cpu_state(bb)
    x = torch.randn(1, 3, 100, 167, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    want = orc.resnet_forward(bsd, x.float(), 18)
    bb = bb.to(cuda_device)
    bb.eval()
    with torch.no_grad():
        feats = bb(x.to(cuda_device))
    _check_levels(feats, want, ["C2", "C3", "C4", "C5"])
    with pytest.raises(RuntimeError):
        neck.to(cuda_device)(feats)


def test_variants(cuda_device):
    """single out index -> bare tensor; RetinaNet-style extra convs with start_level=1."""
    from torch_detection_b200.models.backbone import ResNet
    from torch_detection_b200.models.necks import FPN
    dev = cuda_device
    torch.manual_seed(4)
    bb1 = ResNet(50, out_indices=(3,))
    bb1.init_weights()
    bb1.eval()
    x = torch.randn(1, 3, 128, 128).to(torch.bfloat16)
    want = orc.resnet_forward(helpers.cpu_state(bb1), x.float(), 50, out_indices=(3,))
    with torch.no_grad():
        got = bb1.to(dev)(x.to(dev))
    assert isinstance(got, torch.Tensor)
    assert orc.rel_l2(got.float(), want) <= BF16_GATE
    bb, _ = helpers.build_product_pair(50, seed=4)
    neck = FPN([256, 512, 1024, 2048], 256, 5, start_level=1, add_extra_convs=True)
    neck.init_weights()
    neck.eval()
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    wf = orc.resnet_forward(bsd, x.float(), 50)
    wp = orc.fpn_forward(nsd, wf, [256, 512, 1024, 2048], 256, 5, start_level=1, add_extra_convs=True)
    feats, outs = _run_product(bb, neck, x, dev)
    assert len(outs) == 5
    _check_levels(outs, wp, ["P3", "P4", "P5", "P6", "P7"])


def test_weight_update_invalidates_packed_operands(cuda_device):
    bb, neck = helpers.build_product_pair(18, seed=0)
    x = torch.randn(1, 3, 64, 64).to(torch.bfloat16)
    f0, _ = _run_product(bb, neck, x, cuda_device)
    with torch.no_grad():
        bb.layer1[0].conv1.weight.mul_(0.5)
    f1, _ = _run_product(bb, neck, x, cuda_device)
    assert not torch.equal(f0[0], f1[0])
    want = orc.resnet_forward(helpers.cpu_state(bb), x.float(), 18)
    _check_levels(f1, want, ["C2", "C3", "C4", "C5"])


def test_cpu_tensor_and_train_mode_bn_are_refused(cuda_device):
    from torch_detection_b200.models.backbone import ResNet
    bb = ResNet(18)
    with pytest.raises(NotImplementedError):
        bb.eval()(torch.randn(1, 3, 64, 64))
    bb = ResNet(18, bn_eval=False).to(cuda_device)
    bb.train()
    with pytest.raises(NotImplementedError):
        bb(torch.randn(1, 3, 64, 64, device=cuda_device))


@pytest.mark.parametrize("depth,hw,batch", [
    (18, (512, 512), 2),
    (34, (608, 1024), 1),
    (50, (512, 512), 3),
    (101, (608, 1024), 2),
    (101, (1024, 1024), 1),
])
def test_config5_sweep_depth_by_size(cuda_device, depth, hw, batch):
    """BASELINE.json config 5 (depth x input size x batch sweep) and config 3's ResNet-101: every FPN level of
    every depth at detector-sized inputs against the CPU oracle (image 0; the other images are covered by
    the batch-independence property)."""
    bb, neck = helpers.build_product_pair(depth, seed=13, bnstats=True)
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    g = torch.Generator().manual_seed(depth)
    xb = torch.randn(batch, 3, hw[0], hw[1], generator=g).to(torch.bfloat16)
    want_f, want_p = orc.resnet_fpn_forward(bsd, nsd, xb[:1].float(), depth)
    feats, outs = _run_product(bb, neck, xb, cuda_device)
    assert all(t.shape[0] == batch for t in feats + outs)
    _check_levels([t[:1] for t in feats], want_f, ["C2", "C3", "C4", "C5"])
    e = _check_levels([t[:1] for t in outs], want_p, ["P2", "P3", "P4", "P5", "P6"])
    print("rel-L2 sweep", depth, hw, batch, e)


@pytest.mark.parametrize("activation", [None, "relu", "relu6"])
def test_pafpn_neck(cuda_device, activation):
    """PAFPN (bottom-up path fused as conv + residual) against the CPU oracle."""
    from torch_detection_b200 import models
    from torch_detection_b200.utils import obj_from_dict
    dev = cuda_device
    bb, _ = helpers.build_product_pair(50, seed=6, bnstats=True)
    torch.manual_seed(6)
    neck = obj_from_dict(dict(type="PAFPN", in_channels=[256, 512, 1024, 2048], out_channels=256, num_outs=5,
                              activation=activation), parent=models.necks)
    neck.init_weights()
    neck.eval()
    if activation == "relu6":
        # bring the pyramid (max ~1400 with these BatchNorm statistics) down to the scale of the clamp, so that
        # ConvModule(activation='relu6') (layers.py:114-119) clips a fraction of the outputs, not nearly all of them
        with torch.no_grad():
            for cm in neck.lateral_convs:
                cm.conv.weight.mul_(0.005)
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    x = torch.randn(2, 3, 160, 224, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16)
    wf = orc.resnet_forward(bsd, x.float(), 50)
    wp = orc.pafpn_forward(nsd, [f.clone() for f in wf], [256, 512, 1024, 2048], 256, 5, activation=activation)
    if activation == "relu6":
        clipped = float((wp[1] == 6.0).float().mean())
        assert 0.005 < clipped < 0.2, clipped
    feats, outs = _run_product(bb, neck, x, dev)
    assert len(outs) == 5 and all(o.dtype == torch.bfloat16 for o in outs)
    e = _check_levels(outs, wp, ["N2", "N3", "N4", "N5", "N6"])
    print("rel-L2 PAFPN", activation, e)
    neck.train()
    if activation == "relu6":
        with pytest.raises(NotImplementedError):   # the backward kernels' mask operand encodes ReLU only
            neck(feats)
    else:
        touts = neck(feats)                        # training forward = the same plan (tests/test_train_gpu.py)
        assert all(t.requires_grad for t in touts) and all(torch.equal(a, b) for a, b in zip(touts, outs))


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
def test_input_transform_normalise_and_pad_in_the_loader(cuda_device, dtype):
    """the data layer's (v - mean)/std and pad-to-size-divisor (reference
    datasets/dataset_transforms.py:29-44) folded into TDET_OP_PREP: raw HWC uint8 (or NCHW fp32) batch in,
    features of the normalised, zero-padded image out."""
    dev = cuda_device
    bb, neck = helpers.build_product_pair(50, seed=3, bnstats=True)
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    means, stds = (123.675, 116.28, 103.53), (58.395, 57.12, 57.375)
    g = torch.Generator().manual_seed(2)
    raw = torch.randint(0, 256, (2, 150, 200, 3), generator=g, dtype=torch.uint8)  # N, H, W, C as decoded
    x = raw.permute(0, 3, 1, 2)                                                     # zero-copy NCHW view
    if dtype == torch.float32:
        x = x.float().contiguous()
    norm = (raw.permute(0, 3, 1, 2).float() - torch.tensor(means).view(1, 3, 1, 1)) / torch.tensor(stds).view(1, 3, 1, 1)
    padded = torch.zeros(2, 3, 160, 224)
    padded[:, :, :150, :200] = norm
    # uint8 in -> bf16 path (the staged image is bf16); fp32 in -> the fp32-I/O (split-precision) path
    ref_in = padded if dtype == torch.float32 else padded.to(torch.bfloat16).float()
    gate = FP32_GATE if dtype == torch.float32 else BF16_GATE
    want_f, want_p = orc.resnet_fpn_forward(bsd, nsd, ref_in, 50)
    bb.set_input_transform(means, stds, size_divisor=32)
    bb, neck = bb.to(dev).eval(), neck.to(dev).eval()
    with torch.no_grad():
        feats = bb(x.to(dev))
        outs = neck(feats)
    torch.cuda.synchronize()
    assert tuple(feats[0].shape) == (2, 256, 40, 56)
    feats = [f.float() for f in feats]
    outs = [o.float() for o in outs]
    _check_levels(feats, want_f, ["C2", "C3", "C4", "C5"], gate)
    _check_levels(outs, want_p, ["P2", "P3", "P4", "P5", "P6"], gate)
    bb.set_input_transform(None)
    with torch.no_grad():
        plain = bb(padded.to(torch.bfloat16).to(dev))
    assert orc.rel_l2(plain[0].float(), feats[0]) <= 5e-3   # the same image without the fused transform


@pytest.mark.parametrize("neck_type", ["FPN", "PAFPN"])
def test_neck_with_folded_batchnorm(cuda_device, neck_type):
    """necks built with normalize=... -- the eval-mode BatchNorm after every neck conv is
    folded into the GEMM epilogue (scale/shift), including under the fused upsample-add."""
    from torch_detection_b200 import models
    from torch_detection_b200.utils import obj_from_dict
    dev = cuda_device
    bb, _ = helpers.build_product_pair(18, seed=12, bnstats=True)
    torch.manual_seed(12)
    neck = obj_from_dict(dict(type=neck_type, in_channels=[64, 128, 256, 512], out_channels=256, num_outs=5,
                              normalize=dict(type="BN")), parent=models.necks)
    neck.init_weights()
    sd = neck.state_dict()
    orc.randomize_bn_stats(sd, generator=torch.Generator().manual_seed(7))
    neck.load_state_dict(sd)
    neck.eval()
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    x = torch.randn(2, 3, 128, 192, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16)
    wf = orc.resnet_forward(bsd, x.float(), 18)
    fwd = orc.fpn_forward if neck_type == "FPN" else orc.pafpn_forward
    wp = fwd(nsd, [f.clone() for f in wf], [64, 128, 256, 512], 256, 5)
    feats, outs = _run_product(bb, neck, x, dev)
    e = _check_levels(outs, wp, ["L2", "L3", "L4", "L5", "L6"])
    print("rel-L2 %s + BN" % neck_type, e)
    neck.train()
    with pytest.raises(NotImplementedError):
        neck(feats)


@pytest.mark.parametrize("base_width,cardinality", [(4, 32), (8, 32)])
def test_resnext_backbone(cuda_device, base_width, cardinality):
    """ResNeXt-50 (grouped 3x3 as a 64-channel band on the dense GEMM kernel) + FPN."""
    from torch_detection_b200 import models
    from torch_detection_b200.utils import obj_from_dict
    dev = cuda_device
    torch.manual_seed(5)
    bb = obj_from_dict(dict(type="ResNeXt", depth=50, base_width=base_width, cardinality=cardinality),
                       parent=models.backbone)
    bb.init_weights()
    sd = bb.state_dict()
    orc.randomize_bn_stats(sd, generator=torch.Generator().manual_seed(3))
    bb.load_state_dict(sd)
    bb.eval()
    _, neck = helpers.build_product_pair(50, seed=5)
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    x = torch.randn(2, 3, 128, 160, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16)
    wf = orc.resnet_forward(bsd, x.float(), 50)
    wp = orc.fpn_forward(nsd, [f.clone() for f in wf], [256, 512, 1024, 2048], 256, 5)
    feats, outs = _run_product(bb, neck, x, dev)
    _check_levels(feats, wf, ["C2", "C3", "C4", "C5"])
    e = _check_levels(outs, wp, ["P2", "P3", "P4", "P5", "P6"])
    print("rel-L2 ResNeXt-50 %dx%dd" % (cardinality, base_width), e)
    # (training: tests/test_train_gpu.py::test_resnext_gradients)


@pytest.mark.parametrize("depth,kwargs,shape", [
    (50, dict(dilations=(1, 1, 2, 4), strides=(1, 2, 1, 1)), (1, 3, 128, 160)),   # dilated C4/C5 (stride-8 output)
    (50, dict(strides=(1, 2, 1, 1)), (2, 3, 96, 128)),
    (18, dict(dilations=(1, 1, 2, 2), strides=(1, 2, 1, 1)), (1, 3, 96, 96)),
    (152, dict(), (1, 3, 128, 128)),
    (50, dict(num_stages=3, strides=(1, 2, 2), dilations=(1, 1, 1), out_indices=(0, 1, 2)), (1, 3, 96, 128)),
])
def test_backbone_constructor_variants(cuda_device, depth, kwargs, shape):
    """Module-level dilations != 1, non-default strides, depth 152 and num_stages < 4 (reference
    models/backbone/resnet.py:186-236) against the CPU oracle."""
    from torch_detection_b200.models.backbone import ResNet
    torch.manual_seed(17)
    bb = ResNet(depth, **kwargs)
    bb.init_weights()
    sd = bb.state_dict()
    orc.randomize_bn_stats(sd, generator=torch.Generator().manual_seed(depth))
    bb.load_state_dict(sd)
    bb.eval()
    bsd = helpers.cpu_state(bb)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
    okw = {k: v for k, v in kwargs.items() if k in ("num_stages", "strides", "dilations", "out_indices")}
    want = orc.resnet_forward(bsd, x.float(), depth, **okw)
    with torch.no_grad():
        got = bb.to(cuda_device)(x.to(cuda_device))
    torch.cuda.synchronize()
    assert len(got) == len(want)
    e = _check_levels(got, want, ["C%d" % (i + 2) for i in range(len(want))])
    print("rel-L2 variant", depth, kwargs, e)


def test_checkpoint_round_trip_on_gpu(cuda_device, tmp_path):
    """save_checkpoint -> init_weights(pretrained=path) on a fresh module (reference
    models/utils/checkpoint.py:67-169, resnet.py:240-243) -> CUDA forward equals the oracle on the saved state."""
    from torch_detection_b200.models.backbone import ResNet
    from torch_detection_b200.models.utils import save_checkpoint
    src, neck = helpers.build_product_pair(50, seed=23, bnstats=True)
    path = str(tmp_path / "epoch_1.pth")
    save_checkpoint(src, path, meta=dict(epoch=1))
    torch.manual_seed(99)
    bb = ResNet(50)                      # different random init: everything must come from the file
    bb.init_weights(pretrained=path)
    bb.eval()
    assert helpers.state_hash(bb.state_dict()) == helpers.state_hash(src.state_dict())
    x = torch.randn(2, 3, 96, 128, generator=torch.Generator().manual_seed(4)).to(torch.bfloat16)
    want_f, want_p = orc.resnet_fpn_forward(helpers.cpu_state(src), helpers.cpu_state(neck), x.float(), 50)
    feats, outs = _run_product(bb, neck, x, cuda_device)
    _check_levels(feats, want_f, ["C2", "C3", "C4", "C5"])
    _check_levels(outs, want_p, ["P2", "P3", "P4", "P5", "P6"])
    # a checkpoint loaded into a module that already ran: the packed operands follow the new weights
    other, _ = helpers.build_product_pair(50, seed=24, bnstats=True)
    path2 = str(tmp_path / "other.pth")
    save_checkpoint(other, path2)
    bb.init_weights(pretrained=path2)
    want2 = orc.resnet_forward(helpers.cpu_state(other), x.float(), 50)
    with torch.no_grad():
        got2 = bb(x.to(cuda_device))
    _check_levels(got2, want2, ["C2", "C3", "C4", "C5"])


def _pretrained_like_state(bb, depth, x, seed):
    """Weight / BatchNorm statistics of a TRAINED ResNet rather than a fresh init: heavy-tailed conv weights (filter
    norms log-normal over two decades, 30 % of the taps pruned), gammas log-normal in [1e-2, 4] with dead channels
    (gamma = 0) and two zero-initialised last-BN gammas, non-zero betas, and running statistics CALIBRATED on a batch
    (as training would leave them), so that the folded per-channel scales gamma / sqrt(var) span ~1e-3 ... 10 while
    the activations stay O(1)."""
    g = torch.Generator().manual_seed(seed)
    sd = bb.state_dict()
    for k, v in sd.items():
        if v.dim() == 4:
            boost = torch.exp(1.2 * torch.randn(v.shape[0], generator=g)).clamp(0.05, 20.0)
            sd[k] = torch.randn(v.shape, generator=g) * v.std() * boost.view(-1, 1, 1, 1) * \
                (torch.rand(v.shape, generator=g) < 0.7)
        elif k.endswith("bias") and v.dim() == 1:
            sd[k] = torch.randn(v.shape, generator=g) * 0.3
        elif k.endswith("weight") and v.dim() == 1:
            gamma = torch.exp(0.9 * torch.randn(v.shape, generator=g)).clamp(1e-2, 4.0)
            gamma[torch.rand(v.shape, generator=g) < 0.05] = 0.0
            if ".bn3." in k and ("layer2.1" in k or "layer3.2" in k):
                gamma.zero_()                                                                  # zero-init residual
            if ".bn3." in k or (".bn2." in k and depth < 50):
                gamma *= 0.5                                                                   # residual branches
            sd[k] = gamma
    # calibration pass: every BatchNorm's running statistics = the statistics of its input on this batch
    real_bn = orc._bn

    def calibrating_bn(state, prefix, t):
        state[prefix + ".running_mean"] = t.mean(dim=(0, 2, 3))
        state[prefix + ".running_var"] = t.var(dim=(0, 2, 3), unbiased=False).clamp_min(1e-6)
        return real_bn(state, prefix, t)

    orc._bn = calibrating_bn
    try:
        with torch.no_grad():
            orc.resnet_forward(sd, x, depth)
    finally:
        orc._bn = real_bn
    bb.load_state_dict(sd)
    scales = torch.cat([(sd[k] / torch.sqrt(sd[k[:-6] + "running_var"] + 1e-5)).abs().flatten()
                        for k in sd if k.endswith(".weight") and sd[k].dim() == 1])
    nz = scales[scales > 0]
    return float(nz.min()), float(nz.max())


def _storage_emulation(bsd, nsd, x, depth, fmt):
    """The oracle with every stored activation rounded to `fmt`: "bf16" (north_star's stated storage precision, bf16
    weights) or "fp16" = fp16 significands with an IDEAL per-tensor power-of-two exponent (fp16 weights)."""
    import math

    def ideal_fp16(t):
        a = float(t.abs().max())
        if a == 0.0:
            return t
        e = math.floor(math.log2(a)) - 14
        return (t * 2.0 ** (-e)).to(torch.float16).float() * 2.0 ** e

    saved = orc._r
    orc._r = saved if fmt == "bf16" else ideal_fp16
    try:
        return orc.resnet_fpn_forward_bf16_emulated(bsd, nsd, x, depth, weight_dtype=torch.bfloat16 if fmt == "bf16"
                                                    else torch.float16)
    finally:
        orc._r = saved


@pytest.mark.parametrize("depth,shape", [(50, (2, 3, 160, 224)), (18, (1, 3, 128, 160))])
def test_pretrained_like_statistics(cuda_device, depth, shape):
    """Trained-network statistics instead of a fresh init.  The internal fp16 block-exponent format takes its exponent
    from a worst-case ||w||_1 bound, which gets loose when per-channel scales spread over orders of magnitude: small
    activations could drift into the fp16 subnormals unnoticed.  Such a state is also ill-conditioned in itself
    (rounding the WEIGHTS alone to fp16 already costs ~1e-2 at C5 of ResNet-50, a plain bf16 pipeline ~1e-1), so the
    gate is relative: on every level the CUDA path is no worse than the oracle with plain-bf16 storage -- the
    precision north_star specifies -- and within 1e-2 wherever that emulation is.  Measured: between the bf16 and the
    ideal-exponent fp16 emulation (the stage-boundary tensors and the convs that read them are bf16 by contract)."""
    bb, neck = helpers.build_product_pair(depth, seed=31)
    x = (torch.randn(*shape, generator=torch.Generator().manual_seed(6)) * 1.3 + 0.2).to(torch.bfloat16)
    lo, hi = _pretrained_like_state(bb, depth, x.float(), 100 + depth)
    assert lo < 2e-2 and hi > 5.0, (lo, hi)     # the folded BN scales really span orders of magnitude
    bb.eval()
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    want_f, want_p = orc.resnet_fpn_forward(bsd, nsd, x.float(), depth)
    assert all(torch.isfinite(t).all() and float(t.abs().max()) > 0 for t in want_f)
    names = ["C2", "C3", "C4", "C5", "P2", "P3", "P4", "P5", "P6"]
    want = list(want_f) + list(want_p)
    emu = {}
    for fmt in ("bf16", "fp16"):
        f, p = _storage_emulation(bsd, nsd, x.float(), depth, fmt)
        emu[fmt] = [orc.rel_l2(a, b) for a, b in zip(list(f) + list(p), want)]
    feats, outs = _run_product(bb, neck, x, cuda_device)
    got = [orc.rel_l2(a.float(), b) for a, b in zip(list(feats) + list(outs), want)]
    assert all(torch.isfinite(t.float()).all() for t in list(feats) + list(outs))
    print("rel-L2 pretrained-like %d (folded BN scales %.1e .. %.1e)" % (depth, lo, hi))
    for n, g, b16, f16 in zip(names, got, emu["bf16"], emu["fp16"]):
        print("   %s  cuda %.2e   bf16-storage oracle %.2e   ideal fp16-exponent oracle %.2e" % (n, g, b16, f16))
    bad = {n: (g, b16) for n, g, b16 in zip(names, got, emu["bf16"]) if not g <= max(1.05 * b16, min(BF16_GATE, 2 * b16))}
    assert not bad, "levels worse than the bf16-storage emulation (cuda, emulation): %s" % bad


def test_fpn_end_level(cuda_device):
    """FPN(start_level, end_level) restricting the pyramid to C3..C4 (fpn.py:26-35)."""
    from torch_detection_b200.models.necks import FPN
    bb, _ = helpers.build_product_pair(50, seed=8, bnstats=True)
    torch.manual_seed(8)
    neck = FPN([256, 512, 1024, 2048], 256, 2, start_level=1, end_level=3)
    neck.init_weights()
    neck.eval()
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    x = torch.randn(2, 3, 128, 160, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
    wf = orc.resnet_forward(bsd, x.float(), 50)
    wp = orc.fpn_forward(nsd, [f.clone() for f in wf], [256, 512, 1024, 2048], 256, 2, start_level=1, end_level=3)
    feats, outs = _run_product(bb, neck, x, cuda_device)
    assert len(outs) == 2
    _check_levels(outs, wp, ["P3", "P4"])


def test_plans_built_under_an_sm_reserve(cuda_device):
    """Multi-GPU training sizes every persistent grid to (SMs - 8) so that NCCL kernels have SMs of their own
    (tdet_set_sm_reserve, training.BucketAllReduce.attach): 140-CTA grids and even CTA-pair grids must compute the
    same result -- bit for bit -- as full-width ones."""
    from torch_detection_b200 import engine
    dev = cuda_device
    bb, neck = helpers.build_product_pair(50, seed=5, bnstats=True)
    x = torch.randn(2, 3, 256, 320, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    want_f, want_p = orc.resnet_fpn_forward(helpers.cpu_state(bb), helpers.cpu_state(neck), x.float(), 50)
    f0, p0 = _run_product(bb, neck, x, dev)
    try:
        engine.set_sm_reserve(dev, 8)
        bb._plans.clear()
        neck._plans.clear()
        f1, p1 = _run_product(bb, neck, x, dev)
        grids = [l["grid"] for l in bb._last_run[0].launch_info() + neck._last_run[0].launch_info() if l["kind"] in (1, 3)]
    finally:
        engine.set_sm_reserve(dev, 0)
        bb._plans.clear()
        neck._plans.clear()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    assert max(grids) <= sms - 8, grids
    _check_levels(f1, want_f, ["C2", "C3", "C4", "C5"])
    _check_levels(p1, want_p, ["P2", "P3", "P4", "P5", "P6"])
    for a, b in zip(f0 + p0, f1 + p1):
        assert torch.equal(a, b)


def test_cuda_graph_replay_equals_eager(cuda_device):
    """graphs.GraphedFeatureExtractor: the whole launch sequence of neck(backbone(x)) captured into one CUDA graph
    replays the plans' own kernels: outputs equal the eager run bit for bit, for every new input."""
    from torch_detection_b200.graphs import GraphedFeatureExtractor
    dev = cuda_device
    bb, neck = helpers.build_product_pair(50, seed=7, bnstats=True)
    bb, neck = bb.to(dev).eval(), neck.to(dev).eval()
    g = torch.Generator().manual_seed(3)
    xs = [torch.randn(1, 3, 256, 320, generator=g).to(torch.bfloat16).to(dev) for _ in range(3)]
    graphed = GraphedFeatureExtractor(bb, neck, xs[0])
    for x in xs:
        with torch.no_grad():
            want = [o.clone() for o in neck(bb(x))]
        got = graphed(x)
        torch.cuda.synchronize()
        for a, b in zip(got, want):
            assert torch.equal(a, b)
    with pytest.raises(ValueError):
        graphed(torch.zeros(2, 3, 256, 320, dtype=torch.bfloat16, device=dev))
    bb.train()
    with pytest.raises(NotImplementedError):
        GraphedFeatureExtractor(bb, neck, xs[0])


def test_groupnorm_backbone_and_neck(cuda_device):
    """use_gn=True (nn.GroupNorm(32, C), models/utils/layers.py:50-54; resnet.py:193-232,
    fpn.py:43-61) -- the golden fixture produced by the unmodified reference, then a larger seeded case against the
    oracle with an odd batch, and the refusal to train through GroupNorm."""
    meta, arrays = helpers.load_golden("r50_gn_fpn_64x96")
    bb, neck = helpers.build_product_gn_pair(meta["depth"], seed=meta["seed"])
    assert helpers.state_hash(bb.state_dict()) == meta["bb_hash"]
    feats, outs = _run_product(bb, neck, arrays["x"].to(torch.bfloat16), cuda_device)
    e1 = _check_levels(feats, [arrays["C%d" % i] for i in range(2, 6)], ["C2", "C3", "C4", "C5"])
    e2 = _check_levels(outs, [arrays["P%d" % i] for i in range(2, 7)], ["P2", "P3", "P4", "P5", "P6"])
    print("groupnorm golden", e1, e2)
    x = torch.randn(3, 3, 128, 192, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16)
    want_f = orc.resnet_forward(helpers.cpu_state(bb), x.float(), 50)
    want_p = orc.fpn_forward(helpers.cpu_state(neck), want_f, [256, 512, 1024, 2048], 256, 5)
    feats, outs = _run_product(bb, neck, x, cuda_device)
    e1 = _check_levels(feats, want_f, ["C2", "C3", "C4", "C5"])
    e2 = _check_levels(outs, want_p, ["P2", "P3", "P4", "P5", "P6"])
    print("groupnorm 128x192", e1, e2)
    # fp32 in -> fp32 out of the same values (no split-precision GroupNorm path)
    feats32 = bb(x.float().to(cuda_device))
    assert all(t.dtype == torch.float32 for t in feats32)
    # (the statistics use no atomics: bit-reproducible)
    assert all(torch.equal(a, b.float()) for a, b in zip(feats32, feats))
    # ... and independent of the batch an image travels in (sharded inference is bit-identical)
    one = bb(x[1:2].to(cuda_device))
    assert all(torch.equal(a[1:2], b) for a, b in zip(feats, one))
    bb.train()
    with pytest.raises(NotImplementedError):
        bb(x.to(cuda_device))


def test_groupnorm_basic_block_and_retinanet_neck(cuda_device):
    """use_gn=True with BasicBlock (gn1 / gn2, resnet.py:29-31) and an add_extra_convs neck from level 1."""
    from torch_detection_b200.models.backbone import ResNet
    from torch_detection_b200.models.necks import FPN
    torch.manual_seed(5)
    bb = ResNet(18, use_gn=True)
    bb.init_weights()
    neck = FPN([64, 128, 256, 512], 256, 5, start_level=1, add_extra_convs=True, normalize=dict(type="GN"), use_gn=True)
    neck.init_weights()
    g = torch.Generator().manual_seed(6)
    for mod in (bb, neck):
        sd = mod.state_dict()
        orc.randomize_gn_affine(sd, generator=g)
        mod.load_state_dict(sd)
    x = torch.randn(2, 3, 128, 128, generator=g).to(torch.bfloat16)
    want_f = orc.resnet_forward(helpers.cpu_state(bb), x.float(), 18)
    want_p = orc.fpn_forward(helpers.cpu_state(neck), want_f, [64, 128, 256, 512], 256, 5, start_level=1,
                             add_extra_convs=True)
    feats, outs = _run_product(bb, neck, x, cuda_device)
    _check_levels(feats, want_f, ["C2", "C3", "C4", "C5"])
    e = _check_levels(outs, want_p, ["P3", "P4", "P5", "P6", "P7"])
    print("groupnorm r18 + extra convs", e)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [
    dict(k=1, stride=1, pad=0, bias=True, norm=False, act=None),       # an FPN lateral
    dict(k=3, stride=1, pad=1, bias=True, norm=False, act=None),       # an FPN output conv
    dict(k=3, stride=2, pad=1, bias=False, norm=True, act="relu"),     # conv + BN + ReLU (layers.py type 3)
    dict(k=3, stride=1, pad=1, bias=True, norm=True, act="relu6"),     # norm and bias together (warned, as the reference)
    dict(k=3, stride=1, pad=1, bias=False, norm=True, act="relu", gn=True),   # conv + GroupNorm(32) + ReLU
    dict(k=1, stride=1, pad=0, bias=True, norm=True, act=None, gn=True),
], ids=["1x1_bias", "3x3_bias", "3x3s2_bn_relu", "3x3_bias_bn_relu6", "3x3_gn_relu", "1x1_bias_gn"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_conv_module_stand_alone_call(cuda_device, cfg, dtype):
    """ConvModule.forward outside a neck (layers.py:121-135): one launch with bias / eval BatchNorm / activation in the
    epilogue, against the same module evaluated by PyTorch in fp32 on the bf16-rounded input and weights."""
    import warnings
    from torch_detection_b200.models.utils.layers import ConvModule
    dev = cuda_device
    torch.manual_seed(11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ConvModule(128, 256, cfg["k"], stride=cfg["stride"], padding=cfg["pad"], bias=cfg["bias"],
                       normalize=dict(type="BN") if cfg["norm"] else None, use_gn=cfg.get("gn", False),
                       activation=cfg["act"])
    if cfg["norm"]:
        g = torch.Generator().manual_seed(3)
        with torch.no_grad():
            m.norm.weight.copy_(0.5 + torch.rand(256, generator=g))
            m.norm.bias.copy_(0.2 * torch.randn(256, generator=g))
            if not cfg.get("gn"):
                m.norm.running_mean.copy_(0.3 * torch.randn(256, generator=g))
                m.norm.running_var.copy_(0.5 + torch.rand(256, generator=g))
    m = m.to(dev).eval()
    x = torch.randn(2, 128, 37, 45, generator=torch.Generator().manual_seed(5)).to(dev).to(dtype)
    with torch.no_grad():
        y = m(x)
        xr = x.to(torch.bfloat16).float()
        ref = torch.nn.functional.conv2d(xr, m.conv.weight.to(torch.bfloat16).float(), m.conv.bias, cfg["stride"],
                                         cfg["pad"])
        if cfg["norm"]:
            ref = m.norm(ref)
        if cfg["act"] == "relu":
            ref = torch.relu(ref)
        elif cfg["act"] == "relu6":
            ref = torch.clamp(ref, 0.0, 6.0)
    assert y.dtype == dtype and tuple(y.shape) == tuple(ref.shape)
    err = orc.rel_l2(y.float(), ref)
    assert err <= (1e-2 if cfg.get("gn") else 5e-3), err
    # what stays refused says so: training, batch-statistics BatchNorm
    m.train()
    with pytest.raises(NotImplementedError):
        m(x)
