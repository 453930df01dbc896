"""Cycle trace of the fused bottleneck tail (CTA 0) at layer1's full size: which wait each role sits in.

    python tools/trace_bottleneck_tail.py [--next 0|1] [--tiles K]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch_detection_b200 import engine  # noqa: E402

NAMES = {1: "P patch requested", 2: "W tap requested", 13: "M tap issued", 10: "M patch landed", 11: "M D1 free", 12: "M tap issue", 20: "M G2 wait z2", 21: "M z2 ready",
         22: "M G2 issue", 60: "E E1 wait D1", 61: "E D1 ready", 62: "E z2 published", 70: "E E2 wait D2", 71: "E D2 ready",
         72: "E res slab A landed", 73: "E res slab B landed", 74: "E out slab A published", 75: "E out slab B published",
         80: "E E3 wait D3", 81: "E D3 ready", 82: "E z1' published"}
for j in range(4):
    NAMES[30 + j] = "M G3 chunk %d issue" % j
    NAMES[40 + j] = "R residual %d requested" % j
    NAMES[50 + j] = "S out slab %d -> store" % j
    NAMES[54 + j] = "S store %d read done" % j

ap = argparse.ArgumentParser()
ap.add_argument("--next", type=int, default=1)
ap.add_argument("--tiles", type=int, default=3, help="print the timeline of this many steady-state tiles")
ap.add_argument("--n", type=int, default=16)
ap.add_argument("--planes", type=int, default=64, help="64: layer1 kernel at 200x336; 128: layer2 kernel at 100x168")
ap.add_argument("--all", action="store_true", help="print every traced role, not only the main ones")
args = ap.parse_args()
dev = torch.device("cuda", 0)
pl = args.planes
n, h, w = (args.n, 200, 336) if pl == 64 else (args.n, 100, 168)
if pl == 128:
    args.next = 0
    NAMES.update({2: "W slot requested", 9: "M G1 begin", 12: "M k-block issue", 13: "M k-block wait W", 22: "M D2 free",
                  23: "M W3 tile landed -> issue", 72: "E residual landed", 74: "E out slab published"})
g = torch.Generator().manual_seed(0)
dt = torch.float16
x = torch.randn(n, pl, h, w, generator=g).to(dev).to(dt).contiguous(memory_format=torch.channels_last)
xres = torch.randn(n, 4 * pl, h, w, generator=g).to(dev).to(dt).contiguous(memory_format=torch.channels_last)
w2 = engine.pack_conv_weight((torch.randn(pl, pl, 3, 3, generator=g) * 0.06).to(dev), dt)
w3 = engine.pack_conv_weight((torch.randn(4 * pl, pl, 1, 1, generator=g) * 0.17).to(dev), dt)
w1 = engine.pack_conv_weight((torch.randn(pl, 4 * pl, 1, 1, generator=g) * 0.09).to(dev), dt)
bns = [(torch.ones(c, device=dev), torch.zeros(c, device=dev)) for c in (pl, 4 * pl, pl)]
y = engine.nhwc_empty(n, h, w, 4 * pl, dev, dt)
y2 = engine.nhwc_empty(n, h, w, pl, dev, dt)
trace = torch.zeros(20 * 2048, dtype=torch.int64, device=dev)
nxt = dict(w=w1, bn=bns[2], y=engine.act_of(y2)) if args.next else None
op = engine.op_bottleneck_tail(engine.act_of(x), w2, engine.act_of(y), engine.act_of(xres), w3, bns[0], bns[1], nxt=nxt)
for rep in range(3):
    op.dw = trace.data_ptr() if rep == 2 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.run_op(op, dev)
    e1.record()
    torch.cuda.synchronize()
    print("run %d: %.1f us" % (rep, e0.elapsed_time(e1) * 1e3))
t = trace.cpu().view(20, 2048)
events = []
for wp in range(20):
    for v in t[wp].tolist():
        if v:
            events.append((v >> 8, v & 255, wp))
events.sort()
t0 = events[0][0]
# steady state: tiles after the 20th G2 issue
g2 = [e for e in events if e[1] == 22]
print("G2 issues: %d; mean period %.0f cycles" % (len(g2), (g2[-1][0] - g2[5][0]) / max(len(g2) - 6, 1)))
i0 = 20 if pl == 64 else 10
per = 1 if pl == 64 else 2   # G2 issues per tile
start, end = g2[i0][0], g2[i0 + per * args.tiles][0]
last = {}
for c, code, wp in events:
    if start <= c <= end and (args.all or wp in (0, 1, 4, 8) or code >= 40 and code < 60):
        print("%8d  (+%5d)  warp %2d  %s" % (c - start, c - last.get(wp, c), wp, NAMES.get(code, code)))
    last[wp] = c
taps = [e[0] for e in events if e[1] == 12 and start <= e[0] <= end]
print("tap issues in window: %d, mean gap %.0f cycles" % (len(taps), (taps[-1] - taps[0]) / max(len(taps) - 1, 1)))
# whole-kernel view of CTA 0: first / last event, time to the first G2 issue, per-tile G2 periods
print("CTA 0 span: %d cycles; first event -> first G2 issue: %d; last G2 issue -> last event: %d" %
      (events[-1][0] - t0, g2[0][0] - t0, events[-1][0] - g2[-1][0]))
print("G2 issue gaps:", [g2[i + 1][0] - g2[i][0] for i in range(len(g2) - 1)])
