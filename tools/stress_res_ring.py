"""Stress: full-size layer1 conv3 (+residual+ReLU) repeated; every run must be bit-identical to the first and
match torch's fp32 conv within bf16 rounding.  Used to check the residual-ring protocol (TDET_RES1_RING)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch_detection_b200 import engine
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
n, h, w, cin, cout = 16, 200, 336, 64, 256
x = torch.randn(n, cin, h, w, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
res = torch.randn(n, cout, h, w, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
wt = (torch.randn(cout, cin, 1, 1, generator=g) * 0.1).to(dev)
wp = engine.pack_conv_weight(wt)
scale = (0.5 + torch.rand(cout, generator=g)).to(dev); shift = (0.3 * torch.randn(cout, generator=g)).to(dev)
y = engine.nhwc_empty(n, h, w, cout, dev)
op = engine.op_conv(engine.act_of(x), wp, engine.act_of(y), 1, 1, 1, 0, 1, scale=scale, shift=shift,
                    residual=engine.act_of(res), relu=True)
engine.run_op(op, dev); torch.cuda.synchronize()
first = y.clone()
ref = F.relu(F.conv2d(x.float(), wt.bfloat16().float()) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res.float())
err = float((first.float() - ref).norm() / ref.norm())
bad = 0
big = torch.empty(1 << 28, dtype=torch.float32, device=dev)
for i in range(40):
    big.normal_() if i % 4 == 0 else None   # perturb memory-system timing
    y.zero_()
    engine.run_op(op, dev); torch.cuda.synchronize()
    if not torch.equal(y, first):
        bad += 1
print("rel-L2 vs torch %.3e ; runs differing from the first: %d / 40" % (err, bad))
