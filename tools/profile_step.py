"""Runs K forward steps of the bench workload and exits (target for ncu captures)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench import make_batch  # noqa: E402
from torch_detection_b200 import models  # noqa: E402
from torch_detection_b200.utils import obj_from_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--depth", type=int, default=50)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
exp = 4 if args.depth >= 50 else 1
bb = obj_from_dict(dict(type="ResNet", depth=args.depth), parent=models.backbone)
bb.init_weights()
neck = obj_from_dict(dict(type="FPN", in_channels=[64 * 2 ** i * exp for i in range(4)],
                          out_channels=256, num_outs=5), parent=models.necks)
neck.init_weights()
bb, neck = bb.to(dev).eval(), neck.to(dev).eval()
x = make_batch(args.batch, 800, 1333, 0, torch.bfloat16).to(dev)
with torch.no_grad():
    for i in range(args.steps):
        if i == args.steps - 1:
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_push("tdet_step")  # ncu --nvtx --nvtx-include "tdet_step/" = one warm step
        outs = neck(bb(x))
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("ok", [tuple(o.shape) for o in outs])
