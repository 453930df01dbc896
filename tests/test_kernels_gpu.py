"""Per-kernel parity on the GPU, through the C ABI (tdet_op_run), against the oracle's ATen ops
(F.conv2d / F.max_pool2d in fp32, TF32 disabled) on identical bf16-rounded operands.

Tolerance: the kernels multiply bf16 operands exactly and accumulate in fp32, so the only error vs
the fp32 reference on the same (already bf16-rounded) operands is accumulation order plus ONE bf16
rounding of the stored output (2^-9 relative): rel-L2 <= 4e-3 is asserted (expected ~2e-3).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 4e-3


def rel_l2(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _nhwc(t):
    """logical NCHW tensor -> dense NHWC bf16 buffer viewed as NCHW channels_last."""
    return t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


IM2COL_CASES = [
    # n, h, w, cin, k, stride, pad, dil
    (2, 10, 12, 64, 3, 1, 1, 1),
    (2, 11, 13, 128, 3, 2, 1, 1),
    (3, 9, 14, 64, 1, 2, 0, 1),
    (2, 12, 12, 64, 3, 1, 2, 2),
    (1, 25, 42, 64, 3, 1, 1, 1),
]


@pytest.mark.parametrize("case", IM2COL_CASES)
def test_im2col_tile_semantics(cuda_device, case):
    from torch_detection_b200 import engine
    n, h, w, cin, k, stride, pad, dil = case
    dev = cuda_device
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, cin, h, w, generator=g).to(dev)
    xb = _nhwc(x)
    ho, wo = engine.conv_out(h, k, stride, pad, dil), engine.conv_out(w, k, stride, pad, dil)
    wgt = torch.zeros(64, k, k, cin, dtype=engine.weight_dtype(), device=dev)
    y = engine.nhwc_empty(n, ho, wo, 64, dev)
    op = engine.op_conv((n, h, w, cin), xb, wgt, y, k, k, stride, pad, dil)
    M = n * ho * wo
    xh = xb.permute(0, 2, 3, 1).float().cpu()  # [n][h][w][c]
    bad = []
    for m0 in range(0, M, 128):
        for (r, s) in {(0, 0), (k - 1, k - 1), (k // 2, 0)}:
            for kc in range(cin // 64):
                tile = engine.debug_im2col_tile(op, m0, r, s, kc, dev).float().cpu()
                for i in range(128):
                    m = m0 + i
                    if m >= M:
                        break
                    q = m % wo
                    p = (m // wo) % ho
                    img = m // (wo * ho)
                    ih = p * stride - pad + r * dil
                    iw = q * stride - pad + s * dil
                    if 0 <= ih < h and 0 <= iw < w:
                        exp = xh[img, ih, iw, kc * 64:(kc + 1) * 64]
                    else:
                        exp = torch.zeros(64)
                    if not torch.equal(tile[i], exp):
                        bad.append((m0, r, s, kc, i))
    assert not bad, "im2col mismatches (m0,r,s,kc,row): %s ... total %d" % (bad[:12], len(bad))


CONV_CASES = [
    # name, n, h, w, cin, cout, k, stride, pad, dil
    ("1x1_64_64", 2, 20, 28, 64, 64, 1, 1, 0, 1),
    ("1x1_256_64", 2, 20, 28, 256, 64, 1, 1, 0, 1),
    ("1x1_64_256", 2, 20, 28, 64, 256, 1, 1, 0, 1),
    ("1x1_512_128", 1, 17, 23, 512, 128, 1, 1, 0, 1),
    ("1x1_1024_2048", 1, 7, 11, 1024, 2048, 1, 1, 0, 1),
    ("1x1s2_256_512", 2, 21, 27, 256, 512, 1, 2, 0, 1),
    ("3x3_64_64", 2, 20, 28, 64, 64, 3, 1, 1, 1),
    ("3x3_256_256", 1, 25, 42, 256, 256, 3, 1, 1, 1),
    ("3x3s2_128_128", 2, 21, 27, 128, 128, 3, 2, 1, 1),
    ("3x3s2_512_512", 1, 13, 21, 512, 512, 3, 2, 1, 1),
    ("3x3d2_64_64", 1, 16, 16, 64, 64, 3, 1, 2, 2),
    ("3x3_many_tiles", 4, 50, 84, 64, 64, 3, 1, 1, 1),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
@pytest.mark.parametrize("epi", ["plain", "bn_relu", "bn_res_relu", "bias"])
def test_conv_op(cuda_device, case, epi):
    from torch_detection_b200 import engine
    _, n, h, w, cin, cout, k, stride, pad, dil = case
    dev = cuda_device
    g = torch.Generator().manual_seed(hash(case[0]) % 1000)
    x = torch.randn(n, cin, h, w, generator=g).to(dev)
    wt = (torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5).to(dev)
    xb = _nhwc(x)
    wp = engine.pack_conv_weight(wt)
    assert torch.equal(wp.permute(0, 3, 1, 2).float(), wt.to(engine.weight_dtype()).float())
    ho, wo = engine.conv_out(h, k, stride, pad, dil), engine.conv_out(w, k, stride, pad, dil)
    y = engine.nhwc_empty(n, ho, wo, cout, dev)
    scale = shift = res = None
    relu = False
    if epi in ("bn_relu", "bn_res_relu"):
        scale = (0.5 + torch.rand(cout, generator=g)).to(dev)
        shift = (0.3 * torch.randn(cout, generator=g)).to(dev)
        relu = True
    if epi == "bias":
        shift = (0.3 * torch.randn(cout, generator=g)).to(dev)
    if epi == "bn_res_relu":
        res = _nhwc(torch.randn(n, cout, ho, wo, generator=g).to(dev))
    op = engine.op_conv((n, h, w, cin), xb, wp, y, k, k, stride, pad, dil, scale=scale, shift=shift,
                        residual=res, relu=relu)
    engine.run_op(op, dev)
    torch.cuda.synchronize()
    ref = F.conv2d(xb.float(), wt.to(engine.weight_dtype()).float(), None, stride, pad, dil)
    if scale is not None:
        ref = ref * scale.view(1, -1, 1, 1)
    if shift is not None:
        ref = ref + shift.view(1, -1, 1, 1)
    if res is not None:
        ref = ref + res.float()
    if relu:
        ref = F.relu(ref)
    err = rel_l2(y.float(), ref)
    print("rel-L2 %s %s %.3e" % (case[0], epi, err))
    assert err <= TOL, "rel-L2 %.3e" % err


def test_conv_upsample_add(cuda_device):
    """FPN lateral: 1x1 + bias + nearest-x2 upsample of the coarser level (fpn.py:92-101)."""
    from torch_detection_b200 import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(7)
    n, h, w, cin, cout = 2, 26, 44, 512, 256
    x = torch.randn(n, cin, h, w, generator=g).to(dev)
    wt = (torch.randn(cout, cin, 1, 1, generator=g) * (1.0 / cin) ** 0.5).to(dev)
    bias = (0.1 * torch.randn(cout, generator=g)).to(dev)
    coarse = _nhwc(torch.randn(n, cout, h // 2, w // 2, generator=g).to(dev))
    xb = _nhwc(x)
    wp = engine.pack_conv_weight(wt)
    y = engine.nhwc_empty(n, h, w, cout, dev)
    op = engine.op_conv((n, h, w, cin), xb, wp, y, 1, 1, 1, 0, 1, shift=bias, coarse=coarse,
                        coarse_hw=(h // 2, w // 2))
    engine.run_op(op, dev)
    torch.cuda.synchronize()
    ref = F.conv2d(xb.float(), wt.to(engine.weight_dtype()).float(), bias)
    ref = ref + F.interpolate(coarse.float(), scale_factor=2, mode="nearest")
    err = rel_l2(y.float(), ref)
    assert err <= TOL, "rel-L2 %.3e" % err


@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 70, 101), (2, 224, 320)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_prep_and_stem(cuda_device, shape, dtype):
    """conv1 7x7/2 + bn1 + relu (resnet.py:254-257) on the staged image."""
    from torch_detection_b200 import engine
    dev = cuda_device
    n, h, w = shape
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, 3, h, w, generator=g).to(dev).to(dtype)
    wt = (torch.randn(64, 3, 7, 7, generator=g) * (2.0 / (64 * 49)) ** 0.5).to(dev)
    scale = (0.5 + torch.rand(64, generator=g)).to(dev)
    shift = (0.3 * torch.randn(64, generator=g)).to(dev)
    ho, wo = engine.conv_out(h, 7, 2, 3), engine.conv_out(w, 7, 2, 3)
    hp, wp_ = 2 * ho + 6, 2 * wo + 16
    staged = torch.empty((n, hp, wp_, 4), dtype=torch.bfloat16, device=dev)
    engine.run_op(engine.op_prep(x, staged, ho, wo), dev)
    torch.cuda.synchronize()
    exp = torch.zeros((n, hp, wp_, 4), dtype=torch.bfloat16, device=dev)
    exp[:, 3:3 + h, 3:3 + w, :3] = x.to(torch.bfloat16).permute(0, 2, 3, 1)
    assert torch.equal(staged, exp), "image staging mismatch"
    wpk = engine.pack_stem_weight(wt)
    y = engine.nhwc_empty(n, ho, wo, 64, dev)
    engine.run_op(engine.op_stem(n, h, w, staged, wpk, y, scale, shift), dev)
    torch.cuda.synchronize()
    ref = F.conv2d(x.to(torch.bfloat16).float(), wt.to(engine.weight_dtype()).float(), None, 2, 3)
    ref = F.relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    err = rel_l2(y.float(), ref)
    assert err <= TOL, "rel-L2 %.3e" % err


@pytest.mark.parametrize("shape", [(2, 64, 32, 48), (1, 64, 35, 51), (2, 128, 9, 7)])
def test_maxpool_and_subsample(cuda_device, shape):
    from torch_detection_b200 import engine
    dev = cuda_device
    n, c, h, w = shape
    g = torch.Generator().manual_seed(5)
    x = _nhwc(torch.randn(n, c, h, w, generator=g).to(dev))
    ho, wo = engine.conv_out(h, 3, 2, 1), engine.conv_out(w, 3, 2, 1)
    y = engine.nhwc_empty(n, ho, wo, c, dev)
    engine.run_op(engine.op_maxpool(n, h, w, c, x, y), dev)
    torch.cuda.synchronize()
    assert torch.equal(y, F.max_pool2d(x, 3, 2, 1))  # bit-exact: max of bf16 values
    hs, ws = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    z = engine.nhwc_empty(n, hs, ws, c, dev)
    engine.run_op(engine.op_subsample(n, h, w, c, x, z), dev)
    torch.cuda.synchronize()
    assert torch.equal(z, F.max_pool2d(x, 1, stride=2))


def test_fold_bn(cuda_device):
    from torch_detection_b200 import engine
    dev = cuda_device
    bn = torch.nn.BatchNorm2d(256).to(dev)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
        bn.running_mean.normal_(0, 0.2)
        bn.running_var.uniform_(0.5, 1.5)
    scale, shift = engine.fold_bn(bn)
    es = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    assert torch.allclose(scale, es, rtol=1e-6, atol=0)
    assert torch.allclose(shift, bn.bias - bn.running_mean * es, rtol=1e-5, atol=1e-7)


def test_unsupported_shapes_fail_loudly(cuda_device):
    from torch_detection_b200 import engine, _C
    dev = cuda_device
    x = engine.nhwc_empty(1, 8, 8, 48, dev)
    wgt = torch.zeros(64, 1, 1, 48, dtype=engine.weight_dtype(), device=dev)
    y = engine.nhwc_empty(1, 8, 8, 64, dev)
    with pytest.raises(_C.TdetError):
        engine.run_op(engine.op_conv((1, 8, 8, 48), x, wgt, y, 1, 1, 1, 0), dev)
