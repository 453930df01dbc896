"""Per-kernel parity of the training path (SURVEY.md 8c gradient protocol, tier 1): dgrad / wgrad /
helpers through the C ABI against fp32 autograd of the same op on identical 16-bit-rounded operands
(input, output gradient, weights, ReLU mask).

Tolerance: products of 16-bit operands are exact, accumulation is fp32; data gradients are stored in
bf16 (one 2^-9 rounding -> rel-L2 <= 4e-3, expected ~2e-3); weight gradients are fp32 sums combined
with fp32 atomics (rel-L2 <= 1e-4 asserted, expected ~1e-6).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _nhwc(t, dtype=torch.bfloat16):
    return t.to(dtype).contiguous(memory_format=torch.channels_last)


DGRAD_CASES = [
    # name, n, h, w, cin, cout, k, pad, dil     (stride 1: the dgrad is a plain conv over g)
    ("1x1_256_64", 2, 20, 28, 256, 64, 1, 0, 1),
    ("1x1_64_256", 2, 20, 28, 64, 256, 1, 0, 1),
    ("1x1_2048_512", 1, 7, 11, 2048, 512, 1, 0, 1),
    ("3x3_64_64", 2, 20, 28, 64, 64, 3, 1, 1),
    ("3x3_256_256", 1, 25, 42, 256, 256, 3, 1, 1),
    ("3x3d2_128_128", 1, 16, 18, 128, 128, 3, 2, 2),
]


@pytest.mark.parametrize("case", DGRAD_CASES, ids=[c[0] for c in DGRAD_CASES])
@pytest.mark.parametrize("epi", ["plain", "mask", "res_mask"])
def test_dgrad_stride1(cuda_device, case, epi):
    from torch_detection_b200 import engine
    name, n, h, w, cin, cout, k, pad, dil = case
    dev = cuda_device
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    wt = (torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5).to(dev)
    scale = (0.5 + torch.rand(cout, generator=g)).to(dev)
    gy = _nhwc(torch.randn(n, cout, h, w, generator=g).to(dev))
    wd = engine.pack_dgrad_weight(wt, scale)
    # the packed operand is exactly bf16(scale * w), rotated and transposed
    w_eff = wd.float().permute(3, 0, 1, 2).flip(2, 3).contiguous()  # [co][ci][r][s]
    assert torch.equal(w_eff, (wt * scale.view(-1, 1, 1, 1)).bfloat16().float())
    dx = engine.nhwc_empty(n, h, w, cin, dev)
    res = mask = None
    if epi in ("mask", "res_mask"):
        mask = _nhwc(F.relu(torch.randn(n, cin, h, w, generator=g)).to(dev))
    if epi == "res_mask":
        res = _nhwc(torch.randn(n, cin, h, w, generator=g).to(dev))
    op = engine.op_conv(engine.act_of(gy), wd, engine.act_of(dx), k, k, 1, dil * (k - 1) - pad, dil,
                        residual=engine.act_of(res) if res is not None else None,
                        mask=engine.act_of(mask) if mask is not None else None)
    engine.run_op(op, dev)
    torch.cuda.synchronize()
    x = torch.zeros(n, cin, h, w, device=dev, requires_grad=True)
    yref = F.conv2d(x, w_eff, None, 1, pad, dil)
    (ref,) = torch.autograd.grad(yref, x, gy.float())
    if res is not None:
        ref = ref + res.float()
    if mask is not None:
        ref = ref * (mask.float() > 0)
    err = rel_l2(dx.float(), ref)
    print("dgrad %s %s rel-L2 %.3e" % (name, epi, err))
    assert err <= 4e-3


@pytest.mark.parametrize("case", [("3x3s2_128", 2, 21, 27, 128, 128), ("3x3s2_512", 1, 25, 42, 512, 512),
                                  ("3x3s2_even", 2, 20, 28, 256, 256)], ids=lambda c: c[0])
def test_dgrad_3x3_stride2(cuda_device, case):
    """dgrad of a stride-2 3x3 conv = zero-insertion upsample of g (TDET_OP_DILATE2) + stride-1 conv."""
    from torch_detection_b200 import engine
    name, n, h, w, cin, cout = case
    dev = cuda_device
    g = torch.Generator().manual_seed(11)
    ho, wo = engine.conv_out(h, 3, 2, 1), engine.conv_out(w, 3, 2, 1)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5).to(dev)
    gy = _nhwc(torch.randn(n, cout, ho, wo, generator=g).to(dev))
    mask = _nhwc(F.relu(torch.randn(n, cin, h, w, generator=g)).to(dev))
    wd = engine.pack_dgrad_weight(wt, None)
    gd = engine.nhwc_empty(n, h, w, cout, dev)
    engine.run_op(engine.op_dilate2(engine.act_of(gy), engine.act_of(gd)), dev)
    dx = engine.nhwc_empty(n, h, w, cin, dev)
    engine.run_op(engine.op_conv(engine.act_of(gd), wd, engine.act_of(dx), 3, 3, 1, 1, 1,
                                 mask=engine.act_of(mask)), dev)
    torch.cuda.synchronize()
    x = torch.zeros(n, cin, h, w, device=dev, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv2d(x, wt.bfloat16().float(), None, 2, 1, 1), x, gy.float())
    ref = ref * (mask.float() > 0)
    err = rel_l2(dx.float(), ref)
    print("dgrad %s rel-L2 %.3e" % (name, err))
    assert err <= 4e-3


@pytest.mark.parametrize("case", [("3x3s2_128", 2, 21, 27, 128, 128), ("3x3s2_512", 1, 25, 42, 512, 512),
                                  ("3x3s2_even", 2, 20, 28, 256, 256), ("3x3s2_64_128", 3, 32, 16, 64, 128)],
                         ids=lambda c: c[0])
@pytest.mark.parametrize("scaled", [False, True], ids=["bf16", "fp16_scaled"])
def test_dgrad_3x3_stride2_parity_classes(cuda_device, case, scaled):
    """The training path's stride-2 dgrad: four parity-class convs over the coarse gradient + TDET_OP_PARITY_MERGE
    (training.BackwardBuilder), against autograd of the strided conv with the same 16-bit-rounded weights."""
    from torch_detection_b200 import engine, training
    name, n, h, w, cin, cout = case
    dev = cuda_device
    g = torch.Generator().manual_seed(13)
    ho, wo = engine.conv_out(h, 3, 2, 1), engine.conv_out(w, 3, 2, 1)
    conv = torch.nn.Conv2d(cin, cout, 3, 2, 1, bias=False).to(dev)
    with torch.no_grad():
        conv.weight.copy_((torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5).to(dev))
    scale = (0.5 + torch.rand(cout, generator=g)).to(dev)
    dtype = torch.float16 if scaled else torch.bfloat16
    gy = _nhwc(torch.randn(n, cout, ho, wo, generator=g).to(dev), dtype)
    mask = _nhwc(F.relu(torch.randn(n, cin, h, w, generator=g)).to(dev))
    meta = engine.MetaArena(16, dev) if scaled else None
    bb = training.BackwardBuilder(dev, engine.OperandCache(), dtype=dtype, meta=meta)
    assert training.S2_DGRAD_PARITY
    g_act = engine.Act(gy, (n, ho, wo, cout), dtype, meta.new() if scaled else None)
    if scaled:
        bb.ops.append(engine.op_amax(g_act, g_act.meta))  # the producer of g would have recorded its |max|
    dx = bb.dgrad("c", conv, scale, g_act, (n, h, w, cin), mask=engine.act_of(mask))
    ops, _ = bb.finalize()
    kinds = [o.kind for o in ops]
    assert kinds.count(17) == 1 and 9 not in kinds, "expected the parity path, got op kinds %r" % (kinds,)
    plan = engine.Plan(ops, [], [gy, mask] + bb.buffers, dev, meta=meta)
    plan.run([])
    torch.cuda.synchronize()
    x = torch.zeros(n, cin, h, w, device=dev, requires_grad=True)
    wq = (conv.weight.detach() * scale.view(-1, 1, 1, 1)).to(dtype).float()
    (ref,) = torch.autograd.grad(F.conv2d(x, wq, None, 2, 1, 1), x, gy.float())
    ref = ref * (mask.float() > 0)
    got = dx.buf[:n * h * w * cin].view(dtype).view(n, h, w, cin).permute(0, 3, 1, 2).float()
    if scaled:
        e_dx = meta.read()[(dx.meta - meta.tensor.data_ptr()) // 8][0]
        got = got * 2.0 ** e_dx
    err = rel_l2(got, ref)
    print("parity dgrad %s rel-L2 %.3e" % (name, err))
    assert err <= 4e-3


@pytest.mark.parametrize("hw", [(20, 28), (25, 42)])
def test_dgrad_shortcut_parity_add(cuda_device, hw):
    """Block-input gradient of a strided bottleneck: dgrad(conv1)(g1) + scatter of the stride-2 1x1
    shortcut's dgrad at even pixels (TDET_FLAG_COARSE_PARITY) + external gradient, ReLU-masked."""
    from torch_detection_b200 import engine
    dev = cuda_device
    h, w = hw
    n, cin, planes, cout = 2, 256, 128, 512
    g = torch.Generator().manual_seed(3)
    hc, wc = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    w1 = (torch.randn(planes, cin, 1, 1, generator=g) * 0.1).to(dev)
    wds = (torch.randn(cout, cin, 1, 1, generator=g) * 0.05).to(dev)
    g1 = _nhwc(torch.randn(n, planes, h, w, generator=g).to(dev))
    gout = _nhwc(torch.randn(n, cout, hc, wc, generator=g).to(dev))
    ext = _nhwc(torch.randn(n, cin, h, w, generator=g).to(dev))
    mask = _nhwc(F.relu(torch.randn(n, cin, h, w, generator=g)).to(dev))
    # shortcut dgrad at coarse resolution
    gd = engine.nhwc_empty(n, hc, wc, cin, dev)
    engine.run_op(engine.op_conv(engine.act_of(gout), engine.pack_dgrad_weight(wds), engine.act_of(gd),
                                 1, 1, 1, 0, 1), dev)
    dx = engine.nhwc_empty(n, h, w, cin, dev)
    engine.run_op(engine.op_conv(engine.act_of(g1), engine.pack_dgrad_weight(w1), engine.act_of(dx), 1, 1, 1, 0, 1,
                                 residual=engine.act_of(ext), coarse=engine.act_of(gd), coarse_parity=True,
                                 mask=engine.act_of(mask)), dev)
    torch.cuda.synchronize()
    x = torch.zeros(n, cin, h, w, device=dev, requires_grad=True)
    y1 = F.conv2d(x, w1.bfloat16().float())
    (r1,) = torch.autograd.grad(y1, x, g1.float())
    x2 = torch.zeros(n, cin, h, w, device=dev, requires_grad=True)
    y2 = F.conv2d(x2, wds.bfloat16().float(), None, 2)
    (r2,) = torch.autograd.grad(y2, x2, gout.float())
    # the shortcut dgrad is rounded to bf16 once before it is merged
    r2 = r2.bfloat16().float()
    ref = (r1 + r2 + ext.float()) * (mask.float() > 0)
    err = rel_l2(dx.float(), ref)
    print("block-input gradient rel-L2 %.3e" % err)
    assert err <= 4e-3


WGRAD_CASES = [
    # name, n, h, w, cin, cout, k, stride, pad, dil
    ("1x1_64_64", 2, 20, 28, 64, 64, 1, 1, 0, 1),
    ("1x1_256_64", 2, 20, 28, 256, 64, 1, 1, 0, 1),
    ("1x1_128_512", 3, 17, 23, 128, 512, 1, 1, 0, 1),
    ("1x1_2048_512", 2, 13, 21, 2048, 512, 1, 1, 0, 1),
    ("1x1s2_256_512", 2, 21, 27, 256, 512, 1, 2, 0, 1),
    ("3x3_64_64", 2, 20, 28, 64, 64, 3, 1, 1, 1),
    ("3x3_256_256", 2, 25, 42, 256, 256, 3, 1, 1, 1),
    ("3x3s2_128_128", 2, 21, 27, 128, 128, 3, 2, 1, 1),
    ("3x3d2_64_128", 1, 16, 18, 64, 128, 3, 1, 2, 2),
    ("3x3_many_pixels", 4, 100, 84, 128, 128, 3, 1, 1, 1),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
@pytest.mark.parametrize("xdtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_wgrad(cuda_device, case, xdtype):
    from torch_detection_b200 import engine
    name, n, h, w, cin, cout, k, stride, pad, dil = case
    dev = cuda_device
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    ho, wo = engine.conv_out(h, k, stride, pad, dil), engine.conv_out(w, k, stride, pad, dil)
    x = _nhwc(torch.randn(n, cin, h, w, generator=g).to(dev), xdtype)
    gy = _nhwc(torch.randn(n, cout, ho, wo, generator=g).to(dev), xdtype)  # one format per MMA
    scale = (0.5 + torch.rand(cout, generator=g)).to(dev)
    acc = torch.zeros(cout, k, k, cin, dtype=torch.float32, device=dev)
    engine.run_op(engine.op_wgrad(engine.act_of(x), engine.act_of(gy), acc, k, k, stride, pad, dil, scale=scale), dev)
    dw = torch.empty(cout, cin, k, k, dtype=torch.float32, device=dev)
    engine.run_op(engine.op_dw_unpack(acc, dw, cout, cin, k, k), dev)
    torch.cuda.synchronize()
    wt = torch.zeros(cout, cin, k, k, device=dev, requires_grad=True)
    yref = F.conv2d(x.float(), wt, None, stride, pad, dil) * scale.view(1, -1, 1, 1)
    (ref,) = torch.autograd.grad(yref, wt, gy.float())
    err = rel_l2(dw, ref)
    print("wgrad %s %s rel-L2 %.3e" % (name, xdtype, err))
    assert err <= 1e-4


def test_wgrad_scaled_input(cuda_device):
    """X stored as fp16 significands * 2^e (tdet_tensor_meta): the exponent is applied in the epilogue."""
    from torch_detection_b200 import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(2)
    n, h, w, cin, cout = 2, 12, 20, 128, 256
    xt = torch.randn(n, cin, h, w, generator=g).to(dev) * 40.0
    e = 3
    xs = _nhwc(xt / 2.0 ** e, torch.float16)
    meta = torch.tensor([[e, 0]], dtype=torch.int32, device=dev)
    gy = _nhwc(torch.randn(n, cout, h, w, generator=g).to(dev), torch.float16)
    acc = torch.zeros(cout, 1, 1, cin, dtype=torch.float32, device=dev)
    engine.run_op(engine.op_wgrad(engine.act_of(xs, meta.data_ptr()), engine.act_of(gy), acc, 1, 1, 1, 0), dev)
    torch.cuda.synchronize()
    ref = torch.einsum("nohw,nihw->oi", gy.float(), xs.float() * 2.0 ** e)
    assert rel_l2(acc.view(cout, cin), ref) <= 1e-4


def test_backward_helpers(cuda_device):
    from torch_detection_b200 import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(9)
    n, c, h, w = 3, 256, 26, 42
    x = _nhwc(torch.randn(n, c, h, w, generator=g).to(dev))
    # column sums (bias gradient)
    db = torch.zeros(c, dtype=torch.float32, device=dev)
    engine.run_op(engine.op_colsum(engine.act_of(x), db), dev)
    assert rel_l2(db, x.float().sum(dim=(0, 2, 3))) <= 1e-5
    # 2x2 sum pool (adjoint of nearest x2 upsample)
    y = engine.nhwc_empty(n, h // 2, w // 2, c, dev)
    engine.run_op(engine.op_sumpool2(engine.act_of(x), engine.act_of(y)), dev)
    ref = (F.avg_pool2d(x.float(), 2) * 4).bfloat16()
    assert torch.equal(y, ref)
    up = torch.zeros(n, c, h // 2, w // 2, device=dev, requires_grad=True)
    (adj,) = torch.autograd.grad(F.interpolate(up, scale_factor=2, mode="nearest"), up, x.float())
    assert torch.equal(y, adj.bfloat16())
    # zero-insertion upsample (adjoint of [::2, ::2]) to an odd and an even size
    for (ho, wo) in ((2 * h, 2 * w), (2 * h - 1, 2 * w - 1)):
        z = engine.nhwc_empty(n, ho, wo, c, dev)
        engine.run_op(engine.op_dilate2(engine.act_of(x), engine.act_of(z)), dev)
        ref = torch.zeros(n, c, ho, wo, dtype=torch.bfloat16, device=dev)
        ref[:, :, ::2, ::2] = x
        assert torch.equal(z, ref)
    # masked add
    r = _nhwc(torch.randn(n, c, h, w, generator=g).to(dev))
    m = _nhwc(F.relu(torch.randn(n, c, h, w, generator=g)).to(dev))
    o = engine.nhwc_empty(n, h, w, c, dev)
    engine.run_op(engine.op_add_mask(engine.act_of(x), engine.act_of(o), residual=engine.act_of(r),
                                     mask=engine.act_of(m)), dev)
    ref = ((x.float() + r.float()) * (m.float() > 0)).bfloat16()
    assert torch.equal(o, ref)
    engine.run_op(engine.op_add_mask(engine.act_of(x), engine.act_of(o), mask=engine.act_of(m)), dev)
    assert torch.equal(o, (x.float() * (m.float() > 0)).bfloat16())
    # accumulator reset
    engine.run_op(engine.op_zero(db), dev)
    assert float(db.abs().max()) == 0.0
