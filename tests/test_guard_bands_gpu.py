"""Out-of-bounds writes, checked without compute-sanitizer (it is closed on this GPU pool): every output tensor of
the kernels added in round 2 -- and of the generic conv on ragged tile shapes -- lives inside a larger buffer whose
guard bands (before and after) are filled with a sentinel pattern; after the launch the guards must be untouched and
the payload fully overwritten (no sentinel left where a result is expected).  Shapes are chosen so that the last tile
of every dimension overhangs the tensor."""
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 8192          # elements (16 KiB) on each side, keeps 1 KiB alignment of the payload
SENTINEL = 0x7FC1     # a bf16 / fp16 NaN payload no kernel produces


class Guarded(object):
    def __init__(self, shape, dtype, dev):
        from torch_detection_b200 import engine
        n, h, w, c = shape
        self.numel = n * h * w * c
        self.raw = torch.full((2 * GUARD + self.numel,), SENTINEL, dtype=torch.int16, device=dev)
        self.act = engine.Act(self.raw.view(torch.bfloat16), shape, dtype, None, offset=GUARD)

    def check(self, what):
        r = self.raw.cpu()
        s = torch.tensor(SENTINEL, dtype=torch.int16)
        assert bool((r[:GUARD] == s).all()), "%s: wrote BEFORE its output" % what
        assert bool((r[GUARD + self.numel:] == s).all()), "%s: wrote PAST its output" % what
        assert not bool((r[GUARD:GUARD + self.numel] == s).any()), "%s: left part of its output unwritten" % what


def _nhwc(t, dtype):
    return t.to(dtype).contiguous(memory_format=torch.channels_last)


CONV_SHAPES = [
    # n, h, w, cin, cout, k, stride, pad: ragged against 128-row tiles, 8x16 patches, CTA pairs
    (1, 7, 9, 64, 64, 3, 1, 1),
    (3, 13, 21, 256, 64, 1, 1, 0),
    (2, 25, 42, 256, 256, 3, 1, 1),
    (1, 50, 84, 128, 128, 3, 1, 1),
    (2, 13, 21, 512, 128, 1, 1, 0),
    (1, 27, 45, 256, 256, 3, 2, 1),
    (2, 9, 11, 1024, 2048, 1, 2, 0),
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_outputs_stay_inside(cuda_device, shape):
    from torch_detection_b200 import engine
    n, h, w, cin, cout, k, stride, pad = shape
    dev = cuda_device
    g = torch.Generator().manual_seed(1)
    x = _nhwc(torch.randn(n, cin, h, w, generator=g).to(dev), torch.bfloat16)
    wp = engine.pack_conv_weight((torch.randn(cout, cin, k, k, generator=g) * 0.05).to(dev))
    ho, wo = engine.conv_out(h, k, stride, pad), engine.conv_out(w, k, stride, pad)
    res = _nhwc(torch.randn(n, cout, ho, wo, generator=g).to(dev), torch.bfloat16)
    for relu6 in (False, True):
        y = Guarded((n, ho, wo, cout), torch.bfloat16, dev)
        engine.run_op(engine.op_conv(engine.act_of(x), wp, y.act, k, k, stride, pad, 1, residual=engine.act_of(res),
                                     relu=True, relu6=relu6), dev)
        torch.cuda.synchronize()
        y.check("conv %s relu6=%s" % (shape, relu6))


@pytest.mark.parametrize("shape", [(1, 9, 13), (2, 25, 42), (3, 8, 16), (1, 50, 84)])
def test_dual_conv_and_tail_outputs_stay_inside(cuda_device, shape):
    from torch_detection_b200 import engine
    n, h, w = shape
    dev = cuda_device
    g = torch.Generator().manual_seed(2)
    z = _nhwc(torch.randn(n, 64, h, w, generator=g).to(dev), torch.bfloat16)
    xin = _nhwc(torch.randn(n, 256, h, w, generator=g).to(dev), torch.bfloat16)
    x64 = _nhwc(torch.randn(n, 64, h, w, generator=g).to(dev), torch.bfloat16)
    w2 = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).to(dev)
    w3 = (torch.randn(256, 64, 1, 1, generator=g) * 0.1).to(dev)
    wsc = (torch.randn(256, 64, 1, 1, generator=g) * 0.1).to(dev)
    ones = torch.ones(256, device=dev)
    bn64 = (torch.ones(64, device=dev), torch.zeros(64, device=dev))
    bn256 = (ones, torch.zeros(256, device=dev))
    # fused tail
    y = Guarded((n, h, w, 256), torch.bfloat16, dev)
    engine.run_op(engine.op_bottleneck_tail(engine.act_of(z), engine.pack_conv_weight(w2), y.act, engine.act_of(xin),
                                            engine.pack_conv_weight(w3), bn64, bn256), dev)
    torch.cuda.synchronize()
    y.check("bottleneck tail %s" % (shape,))
    # ... with the next block's conv1
    y = Guarded((n, h, w, 256), torch.bfloat16, dev)
    y2 = Guarded((n, h, w, 64), torch.bfloat16, dev)
    w1n = (torch.randn(64, 256, 1, 1, generator=g) * 0.05).to(dev)
    engine.run_op(engine.op_bottleneck_tail(engine.act_of(z), engine.pack_conv_weight(w2), y.act, engine.act_of(xin),
                                            engine.pack_conv_weight(w3), bn64, bn256,
                                            nxt=dict(w=engine.pack_conv_weight(w1n), bn=bn64, y=y2.act)), dev)
    torch.cuda.synchronize()
    y.check("bottleneck tail + next conv1 %s (y)" % (shape,))
    y2.check("bottleneck tail + next conv1 %s (y2)" % (shape,))
    # dual-source conv3 (projection shortcut in the same launch), stride 1 and 2
    for stride2 in (1, 2):
        hs, ws = (h - 1) * stride2 + 1, (w - 1) * stride2 + 1
        xs = _nhwc(torch.randn(n, 64, hs, ws, generator=g).to(dev), torch.bfloat16)
        wd = engine.pack_dual_weight(w3, ones, wsc, ones)
        y = Guarded((n, h, w, 256), torch.bfloat16, dev)
        engine.run_op(engine.op_conv(engine.act_of(x64), wd, y.act, 1, 1, 1, 0, 1, shift=torch.zeros(256, device=dev),
                                     relu=True, dual=(engine.act_of(xs), stride2)), dev)
        torch.cuda.synchronize()
        y.check("dual conv %s stride2=%d" % (shape, stride2))


@pytest.mark.parametrize("shape", [(1, 7, 5, 64), (2, 13, 21, 256), (3, 2, 3, 2048), (2, 50, 84, 128)])
def test_group_norm_outputs_stay_inside(cuda_device, shape):
    from torch_detection_b200 import engine, _C
    n, h, w, c = shape
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    x = _nhwc(torch.randn(n, c, h, w, generator=g).to(dev), torch.float16)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    numel = engine.gn_stats_numel(n, 32)
    raw = torch.full((2 * 4096 + numel,), float("nan"), dtype=torch.float32, device=dev)
    stats = raw[4096:4096 + numel]
    engine.run_op(engine.op_gn_stats(engine.act_of(x), stats, 32), dev)
    y = Guarded((n, h, w, c), torch.bfloat16, dev)
    engine.run_op(engine.op_gn_apply(engine.act_of(x), stats, 32, gamma, beta, 1e-5, y.act, relu=True), dev)
    torch.cuda.synchronize()
    y.check("gn_apply %s" % (shape,))
    r = raw.cpu()
    assert bool(torch.isnan(r[:4096]).all()) and bool(torch.isnan(r[4096 + numel:]).all()), "gn_stats wrote outside"
    rows = min(_C.GN_STAT_BLOCKS, (h * w * c // 8 + 255) // 256)
    body = r[4096:4096 + numel].view(n, _C.GN_STAT_BLOCKS, 32, 2)
    assert not bool(torch.isnan(body[:, :rows]).any())
    assert rows == _C.GN_STAT_BLOCKS or bool(torch.isnan(body[:, rows:]).all())
