"""fp32-I/O mode (north_star: every FPN level within rel-L2 <= 1e-4 of the fp32 reference): split-precision
kernels.  Every tensor is a bf16 pair value = hi + lo; the GEMM accumulates hi*hi + lo*hi + hi*lo in fp32
(measured 3.5e-5 end to end, TF32 would give 1.2e-3).

Per-kernel gate: rel-L2 <= 2e-5 against torch's fp32 conv on the SAME fp32 operands (expected ~5e-6: the dropped
lo*lo term and the 2^-17 relative residue of each hi+lo pair)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def split_nhwc(t):
    """fp32 logical NCHW -> bf16 (n, 2c, h, w) channels_last: channels [0, c) = hi, [c, 2c) = lo."""
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=1).contiguous(memory_format=torch.channels_last)


def act(t2, c):
    from torch_detection_b200 import engine
    n, _, h, w = t2.shape
    return engine.Act(t2, (n, h, w, c), torch.bfloat16)


def combine(t2):
    c = t2.shape[1] // 2
    return t2[:, :c].float() + t2[:, c:].float()


CASES = [
    # name, n, h, w, cin, cout, k, stride, pad
    ("1x1_64_64", 2, 20, 28, 64, 64, 1, 1, 0),
    ("1x1_256_128", 2, 20, 28, 256, 128, 1, 1, 0),
    ("1x1_128_512", 3, 17, 23, 128, 512, 1, 1, 0),
    ("1x1s2_256_512", 2, 21, 27, 256, 512, 1, 2, 0),
    ("3x3_64_64", 2, 20, 28, 64, 64, 3, 1, 1),
    ("3x3_256_256", 1, 25, 42, 256, 256, 3, 1, 1),
    ("3x3s2_128_128", 2, 21, 27, 128, 128, 3, 2, 1),
    ("1x1_res_many_tiles", 3, 100, 84, 64, 256, 1, 1, 0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("epi", ["plain", "bn_relu", "bn_res_relu"])
def test_split_conv(cuda_device, case, epi):
    from torch_detection_b200 import engine
    name, n, h, w, cin, cout, k, stride, pad = case
    dev = cuda_device
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = torch.randn(n, cin, h, w, generator=g).to(dev)
    wt = (torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5).to(dev)
    wp = engine.pack_conv_weight_split(wt)
    assert torch.equal(wp[..., :cin].float() + wp[..., cin:].float(),
                       (wt.bfloat16().float() + (wt - wt.bfloat16().float()).bfloat16().float()).permute(0, 2, 3, 1))
    ho, wo = engine.conv_out(h, k, stride, pad), engine.conv_out(w, k, stride, pad)
    xs = split_nhwc(x)
    y = torch.empty((n, 2 * cout, ho, wo), dtype=torch.bfloat16, device=dev).contiguous(memory_format=torch.channels_last)
    scale = shift = res = rs = None
    relu = False
    if epi != "plain":
        scale = (0.5 + torch.rand(cout, generator=g)).to(dev)
        shift = (0.3 * torch.randn(cout, generator=g)).to(dev)
        relu = True
    if epi == "bn_res_relu":
        res = torch.randn(n, cout, ho, wo, generator=g).to(dev)
        rs = split_nhwc(res)
    op = engine.op_conv(act(xs, cin), wp, act(y, cout), k, k, stride, pad, 1, scale=scale, shift=shift,
                        residual=act(rs, cout) if rs is not None else None, relu=relu, split=True)
    engine.run_op(op, dev)
    torch.cuda.synchronize()
    ref = F.conv2d(x, wt, None, stride, pad)
    if scale is not None:
        ref = ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if res is not None:
        ref = ref + res
    if relu:
        ref = F.relu(ref)
    err = rel_l2(combine(y), ref)
    print("split conv %s %s rel-L2 %.2e" % (name, epi, err))
    assert err <= 2e-5


def test_split_lateral_with_upsample_add_and_combine(cuda_device):
    from torch_detection_b200 import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(4)
    n, h, w, cin, cout = 2, 24, 40, 512, 256
    x = torch.randn(n, cin, h, w, generator=g).to(dev)
    co = torch.randn(n, cout, h // 2, w // 2, generator=g).to(dev)
    wt = (torch.randn(cout, cin, 1, 1, generator=g) * 0.05).to(dev)
    bias = (0.3 * torch.randn(cout, generator=g)).to(dev)
    y = torch.empty((n, 2 * cout, h, w), dtype=torch.bfloat16, device=dev).contiguous(memory_format=torch.channels_last)
    cs = split_nhwc(co)
    op = engine.op_conv(act(split_nhwc(x), cin), engine.pack_conv_weight_split(wt), act(y, cout), 1, 1, 1, 0, 1,
                        shift=bias, coarse=act(cs, cout), split=True)
    engine.run_op(op, dev)
    out = torch.empty((n, cout, h, w), dtype=torch.float32, device=dev).contiguous(memory_format=torch.channels_last)
    engine.run_op(engine.op_split_combine(act(y, cout), out), dev)
    torch.cuda.synchronize()
    ref = F.conv2d(x, wt, bias) + F.interpolate(co, scale_factor=2, mode="nearest")
    assert torch.equal(out, combine(y))
    assert rel_l2(out, ref) <= 2e-5


@pytest.mark.parametrize("depth,shape,bnstats", [
    (18, (2, 3, 128, 160), True),
    (50, (2, 3, 128, 160), False),
    (50, (1, 3, 224, 320), True),
    (101, (1, 3, 128, 128), True),
])
def test_fp32_io_end_to_end(cuda_device, depth, shape, bnstats):
    """north_star: with fp32 I/O every FPN level (and C2..C5) within rel-L2 <= 1e-4 of the fp32 reference."""
    from oracle import resnet_fpn_oracle as orc
    from tests import helpers
    bb, neck = helpers.build_product_pair(depth, seed=11, bnstats=bnstats)
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(5))
    want_f, want_p = orc.resnet_fpn_forward(bsd, nsd, x, depth)
    bb, neck = bb.to(cuda_device).eval(), neck.to(cuda_device).eval()
    with torch.no_grad():
        feats = bb(x.to(cuda_device))
        outs = neck(feats)
    torch.cuda.synchronize()
    assert all(t.dtype == torch.float32 for t in tuple(feats) + tuple(outs))
    errs = [orc.rel_l2(a, b) for a, b in zip(tuple(feats) + tuple(outs), tuple(want_f) + tuple(want_p))]
    print("fp32 I/O R%d %s rel-L2: C %s  P %s" % (depth, shape, ["%.1e" % e for e in errs[:4]], ["%.1e" % e for e in errs[4:]]))
    assert max(errs) <= 1e-4, errs


def test_fp32_io_full_size_image(cuda_device):
    """BASELINE geometry in fp32-I/O mode: one 800x1333 image zero-padded to 800x1344."""
    from oracle import resnet_fpn_oracle as orc
    from tests import helpers
    bb, neck = helpers.build_product_pair(50, seed=0)
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    x = torch.zeros(1, 3, 800, 1344)
    x[:, :, :, :1333] = torch.randn(1, 3, 800, 1333, generator=torch.Generator().manual_seed(0))
    want_f, want_p = orc.resnet_fpn_forward(bsd, nsd, x, 50)
    bb, neck = bb.to(cuda_device).eval(), neck.to(cuda_device).eval()
    with torch.no_grad():
        outs = neck(bb(x.to(cuda_device)))
    torch.cuda.synchronize()
    errs = [orc.rel_l2(a, b) for a, b in zip(outs, want_p)]
    print("fp32 I/O full size rel-L2", ["%.1e" % e for e in errs])
    assert max(errs) <= 1e-4
