"""BASELINE.json config 5: ResNet-18/34/50/101 + FPN, input 512x512 .. 1344x1344, batch 1-64 on one B200: img/s,
TFLOP/s and (small batches) the latency of the eager launch sequence next to its CUDA-graph replay.

    python tools/sweep_config5.py --out gpurun_out/config5.json [--quick]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench import build_pair  # noqa: E402
from oracle import resnet_fpn_oracle as orc  # noqa: E402  (FLOP model only)
from torch_detection_b200.graphs import GraphedFeatureExtractor  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/config5.json")
ap.add_argument("--quick", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
depths = [18, 34, 50, 101]
sizes = [(512, 512), (800, 1344), (1344, 1344)]
batches = [1, 4, 16, 64]
if args.quick:
    depths, sizes, batches = [50], [(512, 512), (800, 1344)], [1, 16]


def timed(fn, min_s=0.25, min_iters=5):
    fn()
    torch.cuda.synchronize()
    iters = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    while iters < min_iters or time.perf_counter() - t0 < min_s:
        fn()
        iters += 1
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rows = []
for depth in depths:
    bb, neck = build_pair(depth, dev)
    for (h, w) in sizes:
        flops = orc.conv_flops(depth, h, w)[0]
        for b in batches:
            if b * h * w > 64 * 800 * 1344:      # keep the activation arenas within a few tens of GB
                continue
            x = torch.randn(b, 3, h, w, device=dev).to(torch.bfloat16)
            with torch.no_grad():
                eager = timed(lambda: neck(bb(x)))
            row = dict(depth=depth, h=h, w=w, batch=b, eager_ms=eager, img_s=b / eager * 1e3,
                       tflops=b * flops / eager / 1e9, launches=bb._last_run[0].num_launches + neck._last_run[0].num_launches)
            if b <= 4:
                g = GraphedFeatureExtractor(bb, neck, x)
                row["graph_ms"] = timed(lambda: g(x))
                row["graph_img_s"] = b / row["graph_ms"] * 1e3
                del g
            rows.append(row)
            print(json.dumps(row), flush=True)
            bb._plans.clear()
            neck._plans.clear()
            del x
            torch.cuda.empty_cache()
    del bb, neck
    torch.cuda.empty_cache()
with open(args.out, "w") as f:
    json.dump(rows, f, indent=1)
