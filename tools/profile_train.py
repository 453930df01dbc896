"""Runs K training steps (config 4: forward + backward, frozen BN / stem / stage 1) and exits (ncu target)."""
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
ap.add_argument("--batch", type=int, default=8)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
bb = obj_from_dict(dict(type="ResNet", depth=50, frozen_stages=1, bn_eval=True, bn_frozen=True),
                   parent=models.backbone)
bb.init_weights()
neck = obj_from_dict(dict(type="FPN", in_channels=[256, 512, 1024, 2048], out_channels=256, num_outs=5),
                     parent=models.necks)
neck.init_weights()
bb, neck = bb.to(dev).train(), neck.to(dev).train()
x = make_batch(args.batch, 800, 1333, 0, torch.bfloat16).to(dev)
g = torch.Generator().manual_seed(7)
grads = None
for i in range(args.steps):
    if i == args.steps - 1 and i > 0:
        torch.cuda.synchronize()
        # start/end (process-wide) range: backward kernels are launched from autograd's own thread
        rng = torch.cuda.nvtx.range_start("tdet_step")  # ncu --nvtx --nvtx-include "tdet_step" = one warm step
    for p in list(bb.parameters()) + list(neck.parameters()):
        p.grad = None
    outs = neck(bb(x))
    if grads is None:
        grads = [(torch.randn(o.shape, generator=g) * 1e-3).to(torch.bfloat16).to(dev).contiguous(
            memory_format=torch.channels_last) for o in outs]
    torch.autograd.backward(list(outs), grads)
torch.cuda.synchronize()
if args.steps > 1:
    torch.cuda.nvtx.range_end(rng)
print("ok")
