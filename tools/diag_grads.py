"""Per-parameter gradient deviation table (GPU vs teacher-forced oracles) for debugging."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import grad_oracle, resnet_fpn_oracle as orc
from tests import helpers

dev = torch.device("cuda", 0)
depth, frozen = 50, 1
bb, neck = helpers.build_product_pair(depth, seed=21, bnstats=True, frozen_stages=frozen, bn_eval=True, bn_frozen=True)
bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
bb, neck = bb.to(dev).train(), neck.to(dev).train()
g = torch.Generator().manual_seed(5)
x = torch.randn(2, 3, 128, 160, generator=g).to(torch.bfloat16)
outs = neck(bb(x.to(dev)))
grads = [torch.randn(o.shape, generator=g).to(torch.bfloat16) for o in outs]
torch.autograd.backward(list(outs), [t.to(dev).contiguous(memory_format=torch.channels_last) for t in grads])
torch.cuda.synchronize()
got_b = {k: p.grad.detach().cpu() for k, p in bb.named_parameters() if p.grad is not None}
sb, sn = bb.saved_activations(), neck.saved_activations()
tb, tn, _, _ = grad_oracle.teacher_forced_grads(bsd, nsd, sb, sn, depth, grads, train_from_stage=frozen, kernel_rounding=True)
xb, xn, _, _ = grad_oracle.teacher_forced_grads(bsd, nsd, sb, sn, depth, grads, train_from_stage=frozen)
for k in reversed(list(tb)):
    print("%-34s kernel-model %.2e  exact %.2e  model-vs-exact %.2e" % (k, orc.rel_l2(got_b[k], tb[k]), orc.rel_l2(got_b[k], xb[k]), orc.rel_l2(tb[k], xb[k])))
