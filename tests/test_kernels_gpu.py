"""Per-kernel parity on the GPU, through the C ABI (tdet_op_run), against the oracle's ATen ops
(F.conv2d / F.max_pool2d in fp32, TF32 disabled) on identical 16-bit-rounded operands.

Tolerance: the kernels multiply 16-bit operands exactly and accumulate in fp32, so the only error
vs the fp32 reference on the same (already rounded) operands is accumulation order plus ONE rounding
of the stored output: bf16 2^-9 relative -> rel-L2 <= 4e-3 asserted (expected ~1.7e-3); fp16 2^-12
relative -> rel-L2 <= 6e-4 asserted (expected ~2e-4).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {torch.bfloat16: 4e-3, torch.float16: 6e-4}


def rel_l2(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _nhwc(t, dtype=torch.bfloat16):
    """logical NCHW tensor -> dense NHWC 16-bit buffer viewed as NCHW channels_last."""
    return t.to(dtype).contiguous(memory_format=torch.channels_last)


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
    wgt = torch.zeros(64, k, k, cin, dtype=torch.bfloat16, device=dev)
    y = engine.nhwc_empty(n, ho, wo, 64, dev)
    op = engine.op_conv(engine.act_of(xb), wgt, engine.act_of(y), k, k, stride, pad, dil)
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
    ("1x1_res_many_tiles", 3, 100, 84, 64, 256, 1, 1, 0, 1),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
@pytest.mark.parametrize("epi", ["plain", "bn_relu", "bn_res_relu", "bias"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_conv_op(cuda_device, case, epi, dtype):
    from torch_detection_b200 import engine
    name, n, h, w, cin, cout, k, stride, pad, dil = case
    dev = cuda_device
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = torch.randn(n, cin, h, w, generator=g).to(dev)
    wt = (torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5).to(dev)
    xb = _nhwc(x, dtype)
    wp = engine.pack_conv_weight(wt, dtype)
    assert torch.equal(wp.permute(0, 3, 1, 2).float(), wt.to(dtype).float())
    ho, wo = engine.conv_out(h, k, stride, pad, dil), engine.conv_out(w, k, stride, pad, dil)
    y = engine.nhwc_empty(n, ho, wo, cout, dev, dtype)
    scale = shift = res = None
    relu = False
    if epi in ("bn_relu", "bn_res_relu"):
        scale = (0.5 + torch.rand(cout, generator=g)).to(dev)
        shift = (0.3 * torch.randn(cout, generator=g)).to(dev)
        relu = True
    if epi == "bias":
        shift = (0.3 * torch.randn(cout, generator=g)).to(dev)
    if epi == "bn_res_relu":
        res = _nhwc(torch.randn(n, cout, ho, wo, generator=g).to(dev), dtype)
    op = engine.op_conv(engine.act_of(xb), wp, engine.act_of(y), k, k, stride, pad, dil, scale=scale,
                        shift=shift, residual=engine.act_of(res) if res is not None else None,
                        relu=relu)
    engine.run_op(op, dev)
    torch.cuda.synchronize()
    ref = F.conv2d(xb.float(), wt.to(dtype).float(), None, stride, pad, dil)
    if scale is not None:
        ref = ref * scale.view(1, -1, 1, 1)
    if shift is not None:
        ref = ref + shift.view(1, -1, 1, 1)
    if res is not None:
        ref = ref + res.float()
    if relu:
        ref = F.relu(ref)
    err = rel_l2(y.float(), ref)
    print("rel-L2 %s %s %s %.3e" % (name, epi, dtype, err))
    assert err <= TOL[dtype], "rel-L2 %.3e" % err


REVERSE_CASES = [
    # name, n, h, w, cin, cout, k, stride, pad -- one per kernel family / loader (see the ids)
    ("1x1_64_256_res_resident_panel", 3, 100, 84, 64, 256, 1, 1, 0),
    ("1x1_256_1024_res_multi_n_tile", 2, 50, 84, 256, 1024, 1, 1, 0),
    ("1x1_1024_256_pair_long_k", 2, 50, 84, 1024, 256, 1, 1, 0),
    ("3x3_256_512_pair_odd_m_tiles", 1, 25, 42, 256, 512, 3, 1, 1),
    ("3x3_256_256_patch_pair", 1, 48, 32, 256, 256, 3, 1, 1),
    ("3x3_64_64_patch", 4, 50, 84, 64, 64, 3, 1, 1),
    ("3x3_128_128_swapped_patch", 3, 32, 40, 128, 128, 3, 1, 1),
    ("1x1_512_128_swapped", 2, 37, 53, 512, 128, 1, 1, 0),
    ("3x3s2_128_128_im2col", 2, 41, 57, 128, 128, 3, 2, 1),
]


@pytest.mark.parametrize("case", REVERSE_CASES, ids=[c[0] for c in REVERSE_CASES])
def test_conv_reverse_tile_order_is_bit_identical(cuda_device, case):
    """TDET_FLAG_REVERSE (the serpentine schedule tdet_plan_create applies to every second conv of a plan) only changes
    the ORDER in which the persistent kernel visits its tiles: outputs and the recorded |max| must not change by a bit,
    for every loader / tile family (more tiles than CTAs would be the full-size case; here ragged and odd counts)."""
    from torch_detection_b200 import engine
    name, n, h, w, cin, cout, k, stride, pad = case
    dev = cuda_device
    dtype = torch.float16
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    xb = _nhwc(torch.randn(n, cin, h, w, generator=g).to(dev), dtype)
    wp = engine.pack_conv_weight((torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5).to(dev), dtype)
    ho, wo = engine.conv_out(h, k, stride, pad, 1), engine.conv_out(w, k, stride, pad, 1)
    scale = (0.5 + torch.rand(cout, generator=g)).to(dev)
    shift = (0.3 * torch.randn(cout, generator=g)).to(dev)
    res = _nhwc(torch.randn(n, cout, ho, wo, generator=g).to(dev), dtype) if "res" in name else None
    outs = []
    for reverse in (False, True):
        y = engine.nhwc_empty(n, ho, wo, cout, dev, dtype)
        y.fill_(float("nan"))
        op = engine.op_conv(engine.act_of(xb), wp, engine.act_of(y), k, k, stride, pad, 1, scale=scale, shift=shift,
                            residual=engine.act_of(res) if res is not None else None, relu=True, reverse=reverse)
        engine.run_op(op, dev)
        torch.cuda.synchronize()
        outs.append(y)
    assert not torch.isnan(outs[1]).any(), "reverse order left tiles unwritten"
    assert torch.equal(outs[0], outs[1])
    ref = F.relu(F.conv2d(xb.float(), wp.permute(0, 3, 1, 2).float(), None, stride, pad) * scale.view(1, -1, 1, 1)
                 + shift.view(1, -1, 1, 1) + (res.float() if res is not None else 0.0))
    assert rel_l2(outs[1].float(), ref) <= TOL[dtype]


PAIR_CASES = [c for c in CONV_CASES if c[5] % 256 == 0] + [
    ("1x1_1024_256_long_k", 2, 50, 84, 1024, 256, 1, 1, 0, 1),     # 66 m-tiles: the default pair selection
    ("3x3_256_256_odd_tiles", 1, 25, 42, 256, 512, 3, 1, 1, 1),    # 9 m-tiles x 2 n-tiles: padding tile in a pair
    ("3x3_128_128_patch", 3, 32, 40, 128, 128, 3, 1, 1, 1),        # halo-patch loader, 15 spatial tiles (odd)
    ("3x3_256_256_patch", 1, 48, 32, 256, 256, 3, 1, 1, 1),
    ("3x3_128_256_patch_ragged", 2, 30, 38, 128, 256, 3, 1, 1, 1),
]


@pytest.mark.parametrize("case", PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
@pytest.mark.parametrize("epi", ["bn_relu", "bn_res_relu"])
def test_conv_op_cta_pairs(cuda_device, case, epi, monkeypatch):
    """Every eligible conv forced onto the CTA-pair kernel (clusters of two, cta_group::2 MMAs)."""
    from torch_detection_b200 import engine
    monkeypatch.setenv("TDET_PAIR", "15")
    monkeypatch.setenv("TDET_SWAP", "0")
    name, n, h, w, cin, cout, k, stride, pad, dil = case
    dev = cuda_device
    xb = engine.nhwc_empty(n, h, w, cin, dev)
    ho, wo = engine.conv_out(h, k, stride, pad, dil), engine.conv_out(w, k, stride, pad, dil)
    y = engine.nhwc_empty(n, ho, wo, cout, dev)
    wp = engine.pack_conv_weight(torch.zeros(cout, cin, k, k, device=dev))
    plan = engine.Plan([engine.op_conv(engine.act_of(xb), wp, engine.act_of(y), k, k, stride, pad, dil)], [],
                       [xb, y, wp], dev)
    info = plan.launch_info()[0]
    if n * ho * wo > 128:
        assert info["variant"] & 8192, "conv was not scheduled on the pair kernel: %r" % (info,)
    test_conv_op(cuda_device, case, epi, torch.bfloat16)


SWAP_CASES = [c for c in CONV_CASES if c[5] in (64, 128)] + [
    ("1x1_256_128_ragged", 3, 37, 29, 256, 128, 1, 1, 0, 1),
    ("3x3_128_128_two_images", 2, 40, 24, 128, 128, 3, 1, 1, 1),
    ("3x3_64_64_halo_patch", 2, 64, 40, 64, 64, 3, 1, 1, 1),           # 8 x 32 spatial tiles, no waste
    ("3x3_128_128_halo_patch_ragged", 1, 60, 38, 128, 128, 3, 1, 1, 1),  # 12 % waste, clipped stores
    ("3x3_256_128_halo_patch", 3, 32, 16, 256, 128, 3, 1, 1, 1),       # four channel chunks per tile
]


@pytest.mark.parametrize("case", SWAP_CASES, ids=[c[0] for c in SWAP_CASES])
@pytest.mark.parametrize("epi", ["plain", "bn_relu", "bias"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_conv_op_swapped_operands(cuda_device, case, epi, dtype, monkeypatch):
    """Every eligible narrow conv forced onto the operand-swapped kernel (weights as the M operand, 256 pixels as N)."""
    from torch_detection_b200 import engine
    monkeypatch.setenv("TDET_SWAP", "3")
    name, n, h, w, cin, cout, k, stride, pad, dil = case
    dev = cuda_device
    xb = engine.nhwc_empty(n, h, w, cin, dev, dtype)
    ho, wo = engine.conv_out(h, k, stride, pad, dil), engine.conv_out(w, k, stride, pad, dil)
    y = engine.nhwc_empty(n, ho, wo, cout, dev, dtype)
    wp = engine.pack_conv_weight(torch.zeros(cout, cin, k, k, device=dev), dtype)
    plan = engine.Plan([engine.op_conv(engine.act_of(xb), wp, engine.act_of(y), k, k, stride, pad, dil)], [],
                       [xb, y, wp], dev)
    info = plan.launch_info()[0]
    assert info["variant"] & 16384, "conv was not scheduled on the operand-swapped kernel: %r" % (info,)
    test_conv_op(cuda_device, case, epi, dtype)


@pytest.mark.parametrize("shape", [(2, 26, 44), (2, 32, 40), (1, 30, 38), (3, 64, 16)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_conv_upsample_add(cuda_device, shape, dtype):
    """FPN lateral: 1x1 + bias + nearest-x2 upsample of the coarser level (fpn.py:92-101).  The shapes
    cover the global-load path (ragged 8x16 tiling) and the TMA-staged coarse box (spatial tiles)."""
    from torch_detection_b200 import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(7)
    (n, h, w), cin, cout = shape, 512, 256
    x = torch.randn(n, cin, h, w, generator=g).to(dev)
    wt = (torch.randn(cout, cin, 1, 1, generator=g) * (1.0 / cin) ** 0.5).to(dev)
    bias = (0.1 * torch.randn(cout, generator=g)).to(dev)
    coarse = _nhwc(torch.randn(n, cout, h // 2, w // 2, generator=g).to(dev), dtype)
    xb = _nhwc(x, dtype)
    wp = engine.pack_conv_weight(wt, dtype)
    y = engine.nhwc_empty(n, h, w, cout, dev, dtype)
    op = engine.op_conv(engine.act_of(xb), wp, engine.act_of(y), 1, 1, 1, 0, 1, shift=bias,
                        coarse=engine.act_of(coarse))
    engine.run_op(op, dev)
    torch.cuda.synchronize()
    ref = F.conv2d(xb.float(), wt.to(dtype).float(), bias)
    ref = ref + F.interpolate(coarse.float(), scale_factor=2, mode="nearest")
    err = rel_l2(y.float(), ref)
    assert err <= TOL[dtype], "rel-L2 %.3e" % err


DUAL_CASES = [
    # name, n, h2, w2, cin (main), cin2 (shortcut input), cout, stride2
    ("layer1.0", 2, 20, 28, 64, 64, 256, 1),
    ("layer2.0", 2, 21, 27, 128, 256, 512, 2),
    ("layer3.0", 1, 26, 42, 256, 512, 1024, 2),
    ("layer4.0_long_k", 2, 25, 42, 512, 1024, 2048, 2),      # 24 k-blocks: the CTA-pair kernel
    ("many_tiles", 3, 100, 84, 64, 64, 256, 1),               # resident weight panel
    ("many_tiles_s2", 2, 101, 83, 128, 256, 512, 2),
]


def _dual_reference(x, x2, w_a, w_b, sc_a, sc_b, sh_a, sh_b, s2, dtype):
    """fp32 reference on the kernel's operands: the BatchNorm scales are folded into the 16-bit weights."""
    wa = (w_a * sc_a.view(-1, 1, 1, 1)).to(dtype).float()
    wb = (w_b * sc_b.view(-1, 1, 1, 1)).to(dtype).float()
    ref = F.conv2d(x.float(), wa) + F.conv2d(x2.float(), wb, None, s2) + (sh_a + sh_b).view(1, -1, 1, 1)
    return F.relu(ref)


@pytest.mark.parametrize("case", DUAL_CASES, ids=[c[0] for c in DUAL_CASES])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_conv_dual_source(cuda_device, case, dtype):
    """TDET_FLAG_DUAL: conv3 + bn3 and the projection shortcut (1x1 / stride s + BN) + add + ReLU of a stage's
    first bottleneck in ONE launch (resnet.py:110-118, :129-136): one GEMM over [x | x2] with the K-concatenated,
    scale-folded weights, against the two fp32 convs."""
    from torch_detection_b200 import engine
    name, n, h2, w2, cin, cin2, cout, s2 = case
    dev = cuda_device
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    ho, wo = (h2 - 1) // s2 + 1, (w2 - 1) // s2 + 1
    x = _nhwc(torch.randn(n, cin, ho, wo, generator=g).to(dev), dtype)
    x2 = _nhwc(torch.randn(n, cin2, h2, w2, generator=g).to(dev), dtype)
    w_a = (torch.randn(cout, cin, 1, 1, generator=g) * (2.0 / cin) ** 0.5).to(dev)
    w_b = (torch.randn(cout, cin2, 1, 1, generator=g) * (2.0 / cin2) ** 0.5).to(dev)
    sc_a, sc_b = ((0.5 + torch.rand(cout, generator=g)).to(dev) for _ in range(2))
    sh_a, sh_b = ((0.3 * torch.randn(cout, generator=g)).to(dev) for _ in range(2))
    wcat = engine.pack_dual_weight(w_a, sc_a, w_b, sc_b, dtype)
    assert torch.equal(wcat[:, :cin].float(), (w_a * sc_a.view(-1, 1, 1, 1)).to(dtype).float().view(cout, cin))
    assert torch.equal(wcat[:, cin:].float(), (w_b * sc_b.view(-1, 1, 1, 1)).to(dtype).float().view(cout, cin2))
    y = engine.nhwc_empty(n, ho, wo, cout, dev, dtype)
    op = engine.op_conv(engine.act_of(x), wcat, engine.act_of(y), 1, 1, 1, 0, 1, shift=sh_a + sh_b, relu=True,
                        dual=(engine.act_of(x2), s2))
    engine.run_op(op, dev)
    torch.cuda.synchronize()
    ref = _dual_reference(x, x2, w_a, w_b, sc_a, sc_b, sh_a, sh_b, s2, dtype)
    err = rel_l2(y.float(), ref)
    print("dual %s %s rel-L2 %.3e" % (name, dtype, err))
    assert err <= TOL[dtype], "rel-L2 %.3e" % err


def test_conv_dual_source_scaled_output(cuda_device):
    """A dual conv writes an fp16 output with a device-chosen exponent: the bound uses the larger |max| of its two
    (plain bf16) inputs, whose magnitudes differ by orders of magnitude here."""
    from torch_detection_b200 import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    n, h2, w2, cin, cin2, cout, s2 = 2, 30, 44, 128, 256, 512, 2
    ho, wo = (h2 - 1) // s2 + 1, (w2 - 1) // s2 + 1
    for mags in ((1.0, 1.0), (3.0e3, 2.0e-2), (1.0e-3, 40.0)):
        arena = engine.MetaArena(3, dev)
        m_x, m_x2, m_y = arena.new(), arena.new(), arena.new()
        x = _nhwc((torch.randn(n, cin, ho, wo, generator=g) * mags[0]).to(dev))
        x2 = _nhwc((torch.randn(n, cin2, h2, w2, generator=g) * mags[1]).to(dev))
        arena.tensor[0, 1] = x.float().abs().max().view(torch.int32)
        arena.tensor[1, 1] = x2.float().abs().max().view(torch.int32)
        w_a = (torch.randn(cout, cin, 1, 1, generator=g) * (2.0 / cin) ** 0.5).to(dev)
        w_b = (torch.randn(cout, cin2, 1, 1, generator=g) * (2.0 / cin2) ** 0.5).to(dev)
        sc_a, sc_b = ((0.5 + torch.rand(cout, generator=g)).to(dev) for _ in range(2))
        sh = (0.3 * mags[0] * torch.randn(cout, generator=g)).to(dev)
        wcat = engine.pack_dual_weight(w_a, sc_a, w_b, sc_b, torch.bfloat16)
        y = engine.Act(torch.empty(n * ho * wo * cout, dtype=torch.float16, device=dev), (n, ho, wo, cout),
                       torch.float16, m_y)
        op = engine.op_conv(engine.Act(x, (n, ho, wo, cin), torch.bfloat16, m_x), wcat, y, 1, 1, 1, 0, 1, shift=sh,
                            relu=True, consts=engine.bound_consts(wcat, None, sh), scaled_out=True,
                            dual=(engine.Act(x2, (n, h2, w2, cin2), torch.bfloat16, m_x2), s2))
        engine.run_op(op, dev)
        torch.cuda.synchronize()
        e_y, amax_y = arena.read()[2]
        y_true = (y.buf.view(n, ho, wo, cout).float() * 2.0 ** e_y).permute(0, 3, 1, 2)
        ref = _dual_reference(x, x2, w_a, w_b, sc_a, sc_b, sh, torch.zeros_like(sh), s2, torch.bfloat16)
        err = rel_l2(y_true, ref)
        print("dual scaled mags %s: e_y=%d rel-L2 %.2e" % (mags, e_y, err))
        assert torch.isfinite(y_true).all() and err <= TOL[torch.float16]
        assert abs(amax_y - float(ref.abs().max())) <= 2e-3 * amax_y
        assert float(y.buf.float().abs().max()) < 2.0 ** 15


def _tail_reference(x, xres, w2, w3, w1n, bn2, bn3, bn1n, dtype, ydtype):
    """fp32 reference of the fused bottleneck tail on the kernel's operands, rounding z2 and y where the kernel does."""
    def bn(t, p):
        return t * p[0].view(1, -1, 1, 1) + p[1].view(1, -1, 1, 1)
    z2 = F.relu(bn(F.conv2d(x.float(), w2.to(dtype).float(), None, 1, 1), bn2)).to(dtype).float()
    y = F.relu(bn(F.conv2d(z2, w3.to(dtype).float()), bn3) + xres.float())
    y2 = None
    if w1n is not None:
        y2 = F.relu(bn(F.conv2d(y.to(ydtype).float(), w1n.to(ydtype).float()), bn1n))
    return y, y2


@pytest.mark.parametrize("shape", [(2, 24, 40), (1, 50, 84), (3, 100, 84), (2, 37, 29)],
                         ids=["2x24x40", "1x50x84", "3x100x84_many_tiles", "2x37x29_ragged"])
@pytest.mark.parametrize("variant", ["next_conv1", "tail_only", "planes128"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_bottleneck_tail(cuda_device, shape, variant, dtype):
    """TDET_OP_BOTTLENECK_TAIL: conv2 3x3 -> conv3 1x1 + residual + ReLU (-> the next block's conv1) of a layer1
    bottleneck (planes = 64) or a layer2 bottleneck (planes = 128, tail only) (resnet.py:101-118) in one kernel, against
    fp32 convs of the same rounded operands."""
    from torch_detection_b200 import engine
    dev = cuda_device
    n, h, w = shape
    fuse_next = variant == "next_conv1"
    pl = 128 if variant == "planes128" else 64
    g = torch.Generator().manual_seed(h * w + n)
    x = _nhwc(torch.randn(n, pl, h, w, generator=g).to(dev), dtype)
    xres = _nhwc(torch.randn(n, 4 * pl, h, w, generator=g).to(dev), dtype)
    w2 = (torch.randn(pl, pl, 3, 3, generator=g) * (2.0 / (9 * pl)) ** 0.5).to(dev)
    w3 = (torch.randn(4 * pl, pl, 1, 1, generator=g) * (2.0 / pl) ** 0.5).to(dev)
    w1n = (torch.randn(pl, 4 * pl, 1, 1, generator=g) * (2.0 / (4 * pl)) ** 0.5).to(dev)
    bns = [((0.5 + torch.rand(c, generator=g)).to(dev), (0.3 * torch.randn(c, generator=g)).to(dev))
           for c in (pl, 4 * pl, pl)]
    y = engine.nhwc_empty(n, h, w, 4 * pl, dev, dtype)
    y2 = engine.nhwc_empty(n, h, w, pl, dev, dtype)
    y2.zero_()
    nxt = dict(w=engine.pack_conv_weight(w1n, dtype), bn=bns[2], y=engine.act_of(y2)) if fuse_next else None
    op = engine.op_bottleneck_tail(engine.act_of(x), engine.pack_conv_weight(w2, dtype), engine.act_of(y),
                                   engine.act_of(xres), engine.pack_conv_weight(w3, dtype), bns[0], bns[1], nxt=nxt)
    engine.run_op(op, dev)
    torch.cuda.synchronize()
    ref_y, ref_y2 = _tail_reference(x, xres, w2, w3, w1n if fuse_next else None, bns[0], bns[1], bns[2], dtype, dtype)
    err = rel_l2(y.float(), ref_y)
    print("bottleneck tail %s %s %s: y rel-L2 %.3e" % (shape, dtype, variant, err))
    assert err <= TOL[dtype], "y rel-L2 %.3e" % err
    if fuse_next:
        # y2 is computed from the kernel's own 16-bit y: reference it on that
        ref2 = F.relu(F.conv2d(y.float(), w1n.to(dtype).float()) * bns[2][0].view(1, -1, 1, 1) + bns[2][1].view(1, -1, 1, 1))
        err2 = rel_l2(y2.float(), ref2)
        assert err2 <= TOL[dtype], "y2 rel-L2 %.3e" % err2
        assert rel_l2(y2.float(), ref_y2) <= 3 * TOL[dtype]
    # planes = 64: the op equals the unfused launches bit for bit (same arithmetic, same k-block order, same rounding
    # points).  planes = 128: the unfused conv2 runs on the operand-swapped kernel, whose k-block order depends on the
    # loader the shape selects (tap-major im2col vs chunk-major halo patch), so fp32 accumulation may round
    # differently: a handful of 1-ulp flips of the 16-bit result are allowed, nothing more.
    z2 = engine.nhwc_empty(n, h, w, pl, dev, dtype)
    yb = engine.nhwc_empty(n, h, w, 4 * pl, dev, dtype)
    engine.run_op(engine.op_conv(engine.act_of(x), engine.pack_conv_weight(w2, dtype), engine.act_of(z2), 3, 3, 1, 1, 1,
                                 scale=bns[0][0], shift=bns[0][1], relu=True), dev)
    engine.run_op(engine.op_conv(engine.act_of(z2), engine.pack_conv_weight(w3, dtype), engine.act_of(yb), 1, 1, 1, 0, 1,
                                 scale=bns[1][0], shift=bns[1][1], residual=engine.act_of(xres), relu=True), dev)
    torch.cuda.synchronize()
    nd = int((y != yb).sum())
    print("  differing elements vs the unfused pair: %d (max |diff| %.3e)" % (nd, float((y.float() - yb.float()).abs().max())))
    if pl == 64:
        assert torch.equal(y, yb), "fused tail differs from conv2 -> conv3 + residual launched separately"
    else:
        assert nd <= 1e-2 * y.numel() and rel_l2(y.float(), yb.float()) <= 0.1 * TOL[dtype]


def test_bottleneck_tail_scaled(cuda_device):
    """Per-tensor exponents through the fused tail: fp16 input with an exponent, fp16 outputs whose exponents the
    kernel derives from the chained bounds; true values against fp32."""
    from torch_detection_b200 import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    n, h, w = 2, 40, 56
    for mag in (1.0, 2.0e3, 3.0e-4):
        arena = engine.MetaArena(4, dev)
        m_x, m_r, m_y, m_y2 = (arena.new() for _ in range(4))

        def scaled(shape, row):
            t = torch.randn(*shape, generator=g).abs() * mag
            e = int(torch.floor(torch.log2(t.abs().max())).item()) - 13
            stored = _nhwc((t * 2.0 ** (-e)).to(dev), torch.float16)
            arena.tensor[row, 0] = e
            arena.tensor[row, 1] = (stored.float().abs().max() * 2.0 ** e).view(torch.int32)
            return stored, e

        x, e_x = scaled((n, 64, h, w), 0)
        xres, e_r = scaled((n, 256, h, w), 1)
        w2 = (torch.randn(64, 64, 3, 3, generator=g) * (2.0 / 576) ** 0.5).to(dev)
        w3 = (torch.randn(256, 64, 1, 1, generator=g) * (2.0 / 64) ** 0.5).to(dev)
        w1n = (torch.randn(64, 256, 1, 1, generator=g) * (2.0 / 256) ** 0.5).to(dev)
        bns = [((0.5 + torch.rand(c, generator=g)).to(dev), (0.3 * mag * torch.randn(c, generator=g)).to(dev))
               for c in (64, 256, 64)]
        w2p, w3p, w1p = (engine.pack_conv_weight(t, torch.float16) for t in (w2, w3, w1n))
        y = engine.Act(torch.empty(n * h * w * 256, dtype=torch.float16, device=dev), (n, h, w, 256), torch.float16, m_y)
        y2 = engine.Act(torch.empty(n * h * w * 64, dtype=torch.float16, device=dev), (n, h, w, 64), torch.float16, m_y2)
        nxt = dict(w=w1p, bn=bns[2], y=y2, consts=engine.bound_consts(w1p, *bns[2]), scaled_out=True)
        op = engine.op_bottleneck_tail(engine.Act(x, (n, h, w, 64), torch.float16, m_x), w2p, y,
                                       engine.Act(xres, (n, h, w, 256), torch.float16, m_r), w3p, bns[0], bns[1],
                                       consts2=engine.bound_consts(w2p, *bns[0]), consts3=engine.bound_consts(w3p, *bns[1]),
                                       scaled_out=True, nxt=nxt)
        engine.run_op(op, dev)
        torch.cuda.synchronize()
        metas = arena.read()
        (e_y, amax_y), (e_y2, amax_y2) = metas[2], metas[3]
        y_true = (y.buf.view(n, h, w, 256).float() * 2.0 ** e_y).permute(0, 3, 1, 2)
        y2_true = (y2.buf.view(n, h, w, 64).float() * 2.0 ** e_y2).permute(0, 3, 1, 2)
        ref_y, _ = _tail_reference(x.float() * 2.0 ** e_x, xres.float() * 2.0 ** e_r, w2, w3, None, bns[0], bns[1], bns[2],
                                   torch.float16, torch.float16)
        ref_y2 = F.relu(F.conv2d(y_true, w1n.half().float()) * bns[2][0].view(1, -1, 1, 1) + bns[2][1].view(1, -1, 1, 1))
        e1, e2 = rel_l2(y_true, ref_y), rel_l2(y2_true, ref_y2)
        print("bottleneck tail scaled mag %g: e_y=%d e_y2=%d rel-L2 %.2e %.2e" % (mag, e_y, e_y2, e1, e2))
        # (z2 is rounded to fp16 inside the kernel with ITS exponent; the reference rounds the true values: allow 2x)
        assert torch.isfinite(y_true).all() and e1 <= 2 * TOL[torch.float16] and e2 <= TOL[torch.float16]
        assert abs(amax_y - float(y_true.abs().max())) <= 1e-3 * amax_y
        assert abs(amax_y2 - float(y2_true.abs().max())) <= 1e-3 * amax_y2
        assert float(y.buf.float().abs().max()) < 2.0 ** 15 and float(y2.buf.float().abs().max()) < 2.0 ** 15


@pytest.mark.parametrize("magnitude", [1.0, 3.0e4, 2.0e-5])
def test_scaled_fp16_chain(cuda_device, magnitude):
    """Per-tensor power-of-two exponents: conv -> (conv + residual) with device-chosen output
    exponents.  Inputs of magnitude 3e4 would overflow plain fp16 after one layer; 2e-5 would sit in
    the subnormals.  True value = stored * 2^e must match fp32 to fp16 rounding."""
    from torch_detection_b200 import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(11)
    n, h, w, c = 2, 24, 40, 128
    x = (torch.randn(n, c, h, w, generator=g) * magnitude).to(dev)
    w1 = (torch.randn(256, c, 3, 3, generator=g) * 0.05).to(dev)
    w2 = (torch.randn(256, 256, 1, 1, generator=g) * 0.08).to(dev)
    sc1 = (0.5 + torch.rand(256, generator=g)).to(dev)
    sh1 = (0.3 * magnitude * torch.randn(256, generator=g)).to(dev)
    arena = engine.MetaArena(3, dev)
    # bf16 input with known amax (exponent 0), as the network input / a returned stage output
    xb = _nhwc(x)
    m_x, m_t, m_y = arena.new(), arena.new(), arena.new()
    staged_amax = xb.float().abs().max()
    arena.tensor[0, 1] = staged_amax.view(torch.int32)
    w1p = engine.pack_conv_weight(w1, torch.bfloat16)
    w2p = engine.pack_conv_weight(w2, torch.float16)
    t = engine.Act(torch.empty(n * h * w * 256, dtype=torch.float16, device=dev), (n, h, w, 256),
                   torch.float16, m_t)
    y = engine.Act(torch.empty(n * h * w * 256, dtype=torch.float16, device=dev), (n, h, w, 256),
                   torch.float16, m_y)
    c1 = engine.bound_consts(w1p, sc1, sh1)
    c2 = engine.bound_consts(w2p, None, None)
    engine.run_op(engine.op_conv(engine.Act(xb, (n, h, w, c), torch.bfloat16, m_x), w1p, t, 3, 3, 1, 1, 1,
                                 scale=sc1, shift=sh1, relu=True, consts=c1, scaled_out=True), dev)
    engine.run_op(engine.op_conv(t, w2p, y, 1, 1, 1, 0, 1, residual=t, relu=True, consts=c2,
                                 scaled_out=True), dev)
    torch.cuda.synchronize()
    metas = arena.read()
    (e_t, amax_t), (e_y, amax_y) = metas[1], metas[2]
    t_true = t.buf.view(n, h, w, 256).float() * 2.0 ** e_t
    y_true = y.buf.view(n, h, w, 256).float() * 2.0 ** e_y
    assert torch.isfinite(t_true).all() and torch.isfinite(y_true).all()
    ref_t = F.relu(F.conv2d(xb.float(), w1.to(torch.bfloat16).float(), None, 1, 1) * sc1.view(1, -1, 1, 1)
                   + sh1.view(1, -1, 1, 1))
    # second layer consumes the stored (fp16-rounded) t, so reference it on the kernel's own t
    t_nchw = t_true.permute(0, 3, 1, 2)
    ref_y = F.relu(F.conv2d(t_nchw, w2.to(torch.float16).float()) + t_nchw)
    e1 = rel_l2(t_nchw, ref_t)
    e2 = rel_l2(y_true.permute(0, 3, 1, 2), ref_y)
    print("scaled chain magnitude %g: e_t=%d e_y=%d rel-L2 %.2e %.2e" % (magnitude, e_t, e_y, e1, e2))
    assert e1 <= TOL[torch.float16] and e2 <= TOL[torch.float16]
    # recorded |max| is the true one, and stored values use the fp16 range without overflowing
    assert abs(amax_t - float(ref_t.abs().max())) <= 2e-3 * amax_t
    assert abs(amax_y - float(ref_y.abs().max())) <= 2e-3 * amax_y
    for buf in (t.buf, y.buf):
        stored_max = float(buf.float().abs().max())
        assert 2.0 ** 4 <= stored_max < 2.0 ** 15, stored_max


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
    hp, wp_ = engine.stem_staging_dims(ho, wo)
    staged = torch.empty((n, hp, wp_, 4), dtype=torch.bfloat16, device=dev)
    arena = engine.MetaArena(1, dev)
    engine.run_op(engine.op_prep(x, staged, ho, wo, y_meta=arena.new()), dev)
    torch.cuda.synchronize()
    exp = torch.zeros((n, hp, wp_, 4), dtype=torch.bfloat16, device=dev)
    exp[:, 3:3 + h, 3:3 + w, :3] = x.to(torch.bfloat16).permute(0, 2, 3, 1)
    assert torch.equal(staged, exp), "image staging mismatch"
    assert arena.read()[0] == (0, float(x.to(torch.bfloat16).float().abs().max()))
    wpk = engine.pack_stem_weight(wt)
    y = engine.nhwc_empty(n, ho, wo, 64, dev)
    engine.run_op(engine.op_stem(n, h, w, staged, wpk, engine.act_of(y), scale, shift), dev)
    torch.cuda.synchronize()
    ref = F.conv2d(x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), None, 2, 3)
    ref = F.relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    err = rel_l2(y.float(), ref)
    assert err <= TOL[torch.bfloat16], "rel-L2 %.3e" % err
    # the same kernel with the 3x3/2 max-pool in its epilogue (TDET_FLAG_POOL): bit-identical to pooling its output
    hq, wq = engine.conv_out(ho, 3, 2, 1), engine.conv_out(wo, 3, 2, 1)
    yp = engine.nhwc_empty(n, hq, wq, 64, dev)
    yp.fill_(-1.0)
    engine.run_op(engine.op_stem(n, h, w, staged, wpk, engine.act_of(yp), scale, shift, pool=True), dev)
    torch.cuda.synchronize()
    assert torch.equal(yp, F.max_pool2d(y, 3, 2, 1)), "fused stem + max-pool differs from pooling the stem output"


@pytest.mark.parametrize("shape", [(3, 40, 500), (2, 99, 250), (1, 18, 14)])
def test_stem_pool_strips(cuda_device, shape):
    """Fused stem + max-pool over several column strips, odd sizes, and ranges that span images."""
    from torch_detection_b200 import engine
    dev = cuda_device
    n, h, w = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(n, 3, h, w, generator=g).to(dev)
    wt = (torch.randn(64, 3, 7, 7, generator=g) * (2.0 / (64 * 49)) ** 0.5).to(dev)
    scale = (0.5 + torch.rand(64, generator=g)).to(dev)
    shift = (0.3 * torch.randn(64, generator=g)).to(dev)
    ho, wo = engine.conv_out(h, 7, 2, 3), engine.conv_out(w, 7, 2, 3)
    hp, wp_ = engine.stem_staging_dims(ho, wo)
    staged = torch.empty((n, hp, wp_, 4), dtype=torch.bfloat16, device=dev)
    engine.run_op(engine.op_prep(x, staged, ho, wo), dev)
    wpk = engine.pack_stem_weight(wt)
    for dtype in (torch.bfloat16, torch.float16):
        y = engine.nhwc_empty(n, ho, wo, 64, dev, dtype)
        engine.run_op(engine.op_stem(n, h, w, staged, wpk, engine.act_of(y), scale, shift), dev)
        hq, wq = engine.conv_out(ho, 3, 2, 1), engine.conv_out(wo, 3, 2, 1)
        # the pooled rows leave by plain global stores (not clipped by a tensor map): guard bands on both sides
        guard = 4096
        flat = torch.full((guard + n * hq * wq * 64 + guard,), -1.0, dtype=dtype, device=dev)
        yp = flat[guard:guard + n * hq * wq * 64].view(n, hq, wq, 64).permute(0, 3, 1, 2)
        engine.run_op(engine.op_stem(n, h, w, staged, wpk, engine.act_of(yp), scale, shift, pool=True), dev)
        torch.cuda.synchronize()
        assert torch.equal(yp, F.max_pool2d(y.float(), 3, 2, 1).to(dtype)), "fused stem + max-pool mismatch (%s)" % dtype
        assert bool((flat[:guard] == -1).all()) and bool((flat[-guard:] == -1).all()), "write outside the pooled tensor"


@pytest.mark.parametrize("shape", [(2, 64, 32, 48), (1, 64, 35, 51), (2, 128, 9, 7)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_maxpool_and_subsample(cuda_device, shape, dtype):
    from torch_detection_b200 import engine
    dev = cuda_device
    n, c, h, w = shape
    g = torch.Generator().manual_seed(5)
    x = _nhwc(torch.randn(n, c, h, w, generator=g).to(dev), dtype)
    ho, wo = engine.conv_out(h, 3, 2, 1), engine.conv_out(w, 3, 2, 1)
    y = engine.nhwc_empty(n, ho, wo, c, dev, dtype)
    engine.run_op(engine.op_maxpool(engine.act_of(x), engine.act_of(y)), dev)
    torch.cuda.synchronize()
    assert torch.equal(y, F.max_pool2d(x.float(), 3, 2, 1).to(dtype))  # bit-exact: max of stored values
    hs, ws = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    z = engine.nhwc_empty(n, hs, ws, c, dev, dtype)
    engine.run_op(engine.op_subsample(engine.act_of(x), engine.act_of(z)), dev)
    torch.cuda.synchronize()
    assert torch.equal(z, x[:, :, ::2, ::2])


def test_fold_bn_and_bound_consts(cuda_device):
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
    w = torch.randn(256, 128, 3, 3, device=dev) * 0.05
    for dt in (torch.bfloat16, torch.float16):
        wp = engine.pack_conv_weight(w, dt)
        c = engine.bound_consts(wp, scale, shift).cpu()
        g = float((wp.float().abs().sum(dim=(1, 2, 3)) * scale.abs()).max())
        assert g <= float(c[0]) <= g * 1.01
        assert float(c[1]) == float(shift.abs().max())


def test_unsupported_shapes_fail_loudly(cuda_device):
    from torch_detection_b200 import engine, _C
    dev = cuda_device
    x = engine.nhwc_empty(1, 8, 8, 48, dev)
    wgt = torch.zeros(64, 1, 1, 48, dtype=torch.bfloat16, device=dev)
    y = engine.nhwc_empty(1, 8, 8, 64, dev)
    with pytest.raises(_C.TdetError):
        engine.run_op(engine.op_conv(engine.act_of(x), wgt, engine.act_of(y), 1, 1, 1, 0), dev)
    wgt16 = torch.zeros(64, 1, 1, 64, dtype=torch.float16, device=dev)
    x64 = engine.nhwc_empty(1, 8, 8, 64, dev)
    with pytest.raises(ValueError):  # mixed bf16 x fp16 operands are illegal on tcgen05
        engine.op_conv(engine.act_of(x64), wgt16, engine.act_of(y), 1, 1, 1, 0)


GN_CASES = [
    # n, h, w, c, groups, x dtype, exponent, y dtype, residual, coarse, relu
    (2, 12, 20, 64, 32, torch.float16, -3, torch.float16, False, False, True),
    (3, 9, 14, 256, 32, torch.float16, 2, torch.bfloat16, True, False, True),
    (2, 10, 16, 256, 32, torch.bfloat16, 0, torch.bfloat16, False, True, False),
    (1, 7, 5, 2048, 32, torch.float16, -1, torch.float16, True, False, True),
    (2, 50, 84, 128, 32, torch.float16, 0, torch.bfloat16, False, False, False),   # more than one block per image
    (2, 6, 6, 512, 64, torch.bfloat16, 0, torch.float16, True, True, True),
]


@pytest.mark.parametrize("case", GN_CASES)
def test_group_norm_stats_and_apply(cuda_device, case):
    """TDET_OP_GN_STATS + TDET_OP_GN_APPLY against F.group_norm (fp32, CPU) on the same stored 16-bit values:
    nn.GroupNorm(32, C) of the reference's use_gn=True modules (models/utils/layers.py:50-54), with the
    residual add / nearest-upsampled coarse add / ReLU that follow it in the backbone block (resnet.py:111-117)
    and the FPN top-down path (fpn.py:98-101)."""
    from torch_detection_b200 import engine
    n, h, w, c, groups, xdt, e, ydt, with_res, with_coarse, relu = case
    dev = cuda_device
    g = torch.Generator().manual_seed(11)
    # per-channel offsets and scales so that group means / variances are far from 0 / 1
    x = torch.randn(n, c, h, w, generator=g) * (0.5 + torch.rand(1, c, 1, 1, generator=g)) + torch.randn(1, c, 1, 1, generator=g)
    xs = _nhwc(x.to(dev), xdt)
    marena = engine.MetaArena(4, dev)
    xm, rm, ym = marena.new(), marena.new(), marena.new()
    marena.tensor[0, 0] = e
    marena.tensor[1, 0] = 1
    x_true = xs.float().cpu() * 2.0 ** e
    gamma = (0.5 + torch.rand(c, generator=g)).to(dev)
    beta = (0.3 * torch.randn(c, generator=g)).to(dev)
    res = _nhwc(torch.randn(n, c, h, w, generator=g).to(dev), torch.float16) if with_res else None
    coarse = None
    if with_coarse:
        assert h % 2 == 0 and w % 2 == 0
        coarse = _nhwc(torch.randn(n, c, h // 2, w // 2, generator=g).to(dev), torch.bfloat16)
    y = engine.nhwc_empty(n, h, w, c, dev, ydt)
    stats = torch.full((engine.gn_stats_numel(n, groups),), float("nan"), dtype=torch.float32, device=dev)
    xa = engine.act_of(xs, xm)
    engine.run_op(engine.op_gn_stats(xa, stats, groups), dev)
    engine.run_op(engine.op_gn_apply(xa, stats, groups, gamma, beta, 1e-5, engine.act_of(y, ym),
                                     residual=engine.act_of(res, rm) if with_res else None,
                                     coarse=engine.act_of(coarse) if with_coarse else None, relu=relu), dev)
    torch.cuda.synchronize()
    from torch_detection_b200 import _C
    rows = min(_C.GN_STAT_BLOCKS, (h * w * c // 8 + 255) // 256)
    raw = stats.view(n, _C.GN_STAT_BLOCKS, groups, 2).cpu()
    assert not torch.isnan(raw[:, :rows]).any() and (rows == _C.GN_STAT_BLOCKS or torch.isnan(raw[:, rows:]).all())
    st = raw[:, :rows].double().sum(1)
    xg = x_true.double().reshape(n, groups, -1)
    assert torch.allclose(st[..., 0], xg.sum(-1), rtol=1e-4, atol=1e-2 * 2.0 ** e)
    assert torch.allclose(st[..., 1], (xg * xg).sum(-1), rtol=1e-4)
    ref = F.group_norm(x_true, groups, gamma.cpu(), beta.cpu(), 1e-5)
    if with_res:
        ref = ref + res.float().cpu() * 2.0
    if with_coarse:
        ref = ref + F.interpolate(coarse.float().cpu(), scale_factor=2, mode="nearest")
    if relu:
        ref = F.relu(ref)
    assert rel_l2(y.float().cpu(), ref) <= TOL[ydt]
    amax = marena.read()[2]
    # bit-reproducible: a second run of both kernels gives the same bytes
    y2 = engine.nhwc_empty(n, h, w, c, dev, ydt)
    stats2 = torch.zeros_like(stats)
    engine.run_op(engine.op_gn_stats(xa, stats2, groups), dev)
    engine.run_op(engine.op_gn_apply(xa, stats2, groups, gamma, beta, 1e-5, engine.act_of(y2),
                                     residual=engine.act_of(res, rm) if with_res else None,
                                     coarse=engine.act_of(coarse) if with_coarse else None, relu=relu), dev)
    torch.cuda.synchronize()
    assert torch.equal(y.view(torch.int16), y2.view(torch.int16))
    # (the recorded maximum is taken before the output rounding: an upper bound within one ulp of the stored one)
    ymax = float(y.float().abs().max())
    assert amax[0] == 0 and ymax * (1 - 2.0 ** -8) <= amax[1] <= ymax * (1 + 2.0 ** -8)


def test_group_norm_rejects_unsupported_widths(cuda_device):
    from torch_detection_b200 import engine, _C
    dev = cuda_device
    x = engine.nhwc_empty(1, 4, 4, 192, dev)
    stats = torch.zeros(engine.gn_stats_numel(1, 32), dtype=torch.float32, device=dev)
    with pytest.raises(_C.TdetError):
        engine.run_op(engine.op_gn_stats(engine.act_of(x), stats, 32), dev)


@pytest.mark.parametrize("case", [("1x1_256_256", 2, 20, 28, 256, 256, 1, 1, 0), ("3x3s2_256_256", 2, 21, 29, 256, 256, 3, 2, 1),
                                  ("3x3_128_128", 2, 12, 20, 128, 128, 3, 1, 1), ("3x3_64_64", 1, 16, 16, 64, 64, 3, 1, 1)])
def test_conv_relu6(cuda_device, case):
    """TDET_FLAG_RELU6 (ConvModule(activation='relu6'), layers.py:114-119) in every epilogue a conv can be routed to
    (256-wide, operand-swapped 128-wide, 64-wide halo-patch), against F.relu6(F.conv2d) with the clamp biting."""
    from torch_detection_b200 import engine, _C
    name, n, h, w, cin, cout, k, stride, pad = case
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    x = _nhwc(torch.randn(n, cin, h, w, generator=g).to(dev))
    wt = (torch.randn(cout, cin, k, k, generator=g) * (4.0 / (cin * k * k) ** 0.5)).to(dev)
    bias = (torch.randn(cout, generator=g)).to(dev)
    wp = engine.pack_conv_weight(wt)
    ho, wo = engine.conv_out(h, k, stride, pad), engine.conv_out(w, k, stride, pad)
    y = engine.nhwc_empty(n, ho, wo, cout, dev)
    engine.run_op(engine.op_conv(engine.act_of(x), wp, engine.act_of(y), k, k, stride, pad, 1, shift=bias, relu6=True), dev)
    torch.cuda.synchronize()
    ref = F.relu6(F.conv2d(x.float().cpu(), wp.permute(0, 3, 1, 2).float().cpu(), bias.cpu(), stride, pad))
    clipped = float((ref == 6.0).float().mean())
    assert 0.01 < clipped < 0.5, clipped
    assert float(y.float().max()) == 6.0 and float(y.float().min()) == 0.0
    assert rel_l2(y.float().cpu(), ref) <= TOL[torch.bfloat16]
    with pytest.raises(_C.TdetError):   # plain outputs only
        meta = engine.MetaArena(2, dev)
        consts = engine.bound_consts(wp, None, bias)
        yh = engine.nhwc_empty(n, ho, wo, cout, dev, torch.float16)
        engine.run_op(engine.op_conv(engine.act_of(x, meta.new()), wp, engine.act_of(yh, meta.new()), k, k, stride, pad, 1,
                                     shift=bias, relu6=True, consts=consts, scaled_out=True), dev)


@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 37, 53)])
@pytest.mark.parametrize("src", [torch.float32, torch.uint8])
def test_stem_fp16_staging(cuda_device, shape, src):
    """TDET_OP_PREP with y_dtype = F16 + the stem over it (fp16 weights): used ahead of GroupNorm chains.  The staged
    values keep 11 significand bits of the fp32 / uint8 input (bf16: 8), the conv matches fp32 on fp16-rounded
    operands to the fp16 tolerance, out-of-range pixels saturate."""
    from torch_detection_b200 import engine
    dev = cuda_device
    n, h, w = shape
    g = torch.Generator().manual_seed(5)
    if src == torch.uint8:
        x = torch.randint(0, 256, (n, 3, h, w), generator=g, dtype=torch.uint8).to(dev)
        sc = torch.tensor([1 / 58.4, 1 / 57.1, 1 / 57.4], device=dev)
        sh = torch.tensor([-123.7 / 58.4, -116.3 / 57.1, -103.5 / 57.4], device=dev)
        want = x.float() * sc.view(1, 3, 1, 1) + sh.view(1, 3, 1, 1)
    else:
        x = torch.randn(n, 3, h, w, generator=g).to(dev)
        x[0, 0, 0, 0] = 1e6     # saturates at 65504 in the staging
        sc = sh = None
        want = x.clamp(-65504, 65504)
    wt = (torch.randn(64, 3, 7, 7, generator=g) * (2.0 / (64 * 49)) ** 0.5).to(dev)
    ho, wo = engine.conv_out(h, 7, 2, 3), engine.conv_out(w, 7, 2, 3)
    hp, wp_ = engine.stem_staging_dims(ho, wo)
    staged = torch.empty((n, hp, wp_, 4), dtype=torch.float16, device=dev)
    arena = engine.MetaArena(2, dev)
    m_in, m_out = arena.new(), arena.new()
    engine.run_op(engine.op_prep(x, staged, ho, wo, y_meta=m_in, scale=sc, shift=sh, y_dtype=torch.float16), dev)
    torch.cuda.synchronize()
    exp = torch.zeros((n, hp, wp_, 4), dtype=torch.float16, device=dev)
    exp[:, 3:3 + h, 3:3 + w, :3] = want.to(torch.float16).permute(0, 2, 3, 1)
    assert torch.equal(staged, exp), "fp16 image staging mismatch"
    assert arena.read()[0] == (0, float(exp.float().abs().max()))
    if src == torch.float32:
        x[0, 0, 0, 0] = 0.5
        engine.run_op(engine.op_prep(x, staged, ho, wo, y_meta=m_in, y_dtype=torch.float16), dev)
        want = x
    wpk = engine.pack_stem_weight(wt, dtype=torch.float16)
    assert wpk.dtype == torch.float16
    consts = engine.bound_consts(wpk, None, None)
    y = engine.nhwc_empty(n, ho, wo, 64, dev, torch.float16)
    engine.run_op(engine.op_stem(n, h, w, staged, wpk, engine.act_of(y, m_out), None, None, relu=False, x_meta=m_in,
                                 consts=consts, scaled_out=True, x_dtype=torch.float16), dev)
    torch.cuda.synchronize()
    e = arena.read()[1][0]
    ref = F.conv2d(want.to(torch.float16).float(), wt.to(torch.float16).float(), None, 2, 3)
    err = rel_l2(y.float() * 2.0 ** e, ref)
    assert err <= TOL[torch.float16], "rel-L2 %.3e" % err
    # mixing formats is refused
    with pytest.raises(Exception):
        engine.run_op(engine.op_prep(x, staged, ho, wo, split=True, y_dtype=torch.float16), dev)
