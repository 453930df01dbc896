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


@pytest.mark.parametrize("activation", [None, "relu"])
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
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    x = torch.randn(2, 3, 160, 224, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16)
    wf = orc.resnet_forward(bsd, x.float(), 50)
    wp = orc.pafpn_forward(nsd, [f.clone() for f in wf], [256, 512, 1024, 2048], 256, 5, activation=activation)
    feats, outs = _run_product(bb, neck, x, dev)
    assert len(outs) == 5 and all(o.dtype == torch.bfloat16 for o in outs)
    e = _check_levels(outs, wp, ["N2", "N3", "N4", "N5", "N6"])
    print("rel-L2 PAFPN", activation, e)
    neck.train()
    with pytest.raises(NotImplementedError):
        neck(feats)


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
    bb.train()
    with pytest.raises(NotImplementedError):
        bb(x.to(dev))
