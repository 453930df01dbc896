"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, on
shapes a sanitizer finishes in minutes -- a tiny ResNet-50 + FPN inference step (stem + fused pool, halo-patch and
im2col convs, dual-source conv3, fused bottleneck tails, operand-swapped kernel, FPN laterals with the TMA-staged coarse
box), the same with CTA pairs forced, one training step (dgrad with mask ring, parity-class stride-2 dgrads,
wgrad, helpers) and a GroupNorm backbone + neck (--gn: raw convs, statistics / apply kernels).

    compute-sanitizer --tool memcheck python tools/sanitize_case.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench import build_pair  # noqa: E402

dev = torch.device("cuda", 0)
x = torch.randn(1, 3, 128, 160, generator=torch.Generator().manual_seed(0)).to(torch.bfloat16).to(dev)
bb, neck = build_pair(50, dev)
with torch.no_grad():
    outs = neck(bb(x))
torch.cuda.synchronize()
print("inference ok", [tuple(o.shape) for o in outs], float(outs[0].float().abs().mean()))
if "--pairs" in sys.argv:
    os.environ["TDET_PAIR"] = "15"
    bb._plans.clear()
    neck._plans.clear()
    with torch.no_grad():
        outs2 = neck(bb(x))
    torch.cuda.synchronize()
    print("forced CTA pairs ok", all(torch.equal(a, b) for a, b in zip(outs, outs2)))
if "--train" in sys.argv:
    bb, neck = build_pair(50, dev, train=True)
    o = neck(bb(x))
    torch.autograd.backward(list(o), [torch.ones_like(t) * 1e-3 for t in o])
    torch.cuda.synchronize()
    print("training step ok", sum(1 for p in list(bb.parameters()) + list(neck.parameters()) if p.grad is not None))
if "--gn" in sys.argv:
    from torch_detection_b200.models.backbone import ResNet
    from torch_detection_b200.models.necks import FPN
    torch.manual_seed(1)
    gbb = ResNet(50, use_gn=True)
    gbb.init_weights()
    gneck = FPN([256, 512, 1024, 2048], 256, 5, normalize=dict(type="GN"), use_gn=True)
    gneck.init_weights()
    gbb, gneck = gbb.to(dev).eval(), gneck.to(dev).eval()
    with torch.no_grad():
        go = gneck(gbb(x))
    torch.cuda.synchronize()
    print("groupnorm inference ok", [tuple(o.shape) for o in go], float(go[0].float().abs().mean()))
