"""CPU checks of the gradient oracle (oracle/grad_oracle.py): the teacher-forced restatement, forced
with the fp32 oracle's OWN activations and bf16-exact weights, must reproduce plain autograd -- so the
only thing the GPU gate can measure is the kernels' arithmetic."""
import pytest
import torch

from oracle import grad_oracle, resnet_fpn_oracle as orc


@pytest.mark.parametrize("depth,frozen", [(18, 0), (18, 1), (50, 1), (50, 2)])
def test_teacher_forcing_with_own_activations_is_plain_autograd(depth, frozen):
    torch.set_num_threads(4)
    g = torch.Generator().manual_seed(4)
    bsd = orc.make_resnet_state(depth, generator=g)
    exp = orc.EXPANSION[orc.ARCH[depth][0]]
    nsd = orc.make_fpn_state([64 * 2 ** i * exp for i in range(4)], 256, 5, generator=g)
    orc.randomize_bn_stats(bsd, generator=g)
    # bf16-exact conv weights: the straight-through weight rounding is then the identity
    bsd = {k: (v.bfloat16().float() if v.dim() == 4 else v) for k, v in bsd.items()}
    nsd = {k: (v.bfloat16().float() if v.dim() == 4 else v) for k, v in nsd.items()}
    x = torch.randn(1, 3, 64, 96, generator=g)
    with torch.no_grad():
        _, outs = orc.resnet_fpn_forward(bsd, nsd, x, depth)
    grads = [torch.randn(o.shape, generator=g) for o in outs]
    pb, pn, _, _ = grad_oracle.plain_grads(bsd, nsd, x, depth, grads, train_from_stage=frozen)
    sb, sn = grad_oracle.oracle_saved_activations(bsd, nsd, x, depth)
    tb, tn, _, touts = grad_oracle.teacher_forced_grads(bsd, nsd, sb, sn, depth, grads, train_from_stage=frozen)
    assert set(tb) == set(pb) and set(tn) == set(pn)
    assert all(k.startswith(("layer%d" % (i + 1)) ) for k in tb for i in [int(k[5]) - 1]) and \
        all(int(k[5]) - 1 >= frozen for k in tb)
    for a, b in zip(outs, touts):
        assert orc.rel_l2(b, a) < 1e-5
    for k in pn:
        assert orc.rel_l2(tn[k], pn[k]) < 1e-4, k
    for k in pb:
        # the shortcut branch is re-rounded to bf16 in the forced graph (straight-through): values move
        # by 2^-9 relative, gradients do not
        assert orc.rel_l2(tb[k], pb[k]) < 1e-4, k
