"""End-to-end gradient parity of the training configuration (BASELINE.json config 4; SURVEY.md 8c
gradient protocol) on the GPU: product ResNet(frozen stem + stage 1, frozen BN) + FPN in train mode,
fixed random upstream gradients on P2..P6, torch.autograd.backward through the modules' hand-written
backward plans, against the CPU gradient oracle.

Gates (bf16 tolerance of north_star, rel-L2 <= 1e-2), all against the TEACHER-FORCED oracle = fp32
autograd over the kernels' own stored activations / ReLU masks (grad_oracle.teacher_forced_grads):
  * EVERY parameter gradient -- the 16 of the FPN and every trainable backbone conv weight (42 for
    ResNet-50 with a frozen stage 1, 80 for ResNet-101 from stage 3) -- vs the exact fp32 backward:
    <= 1e-2.  Measured 4.1-6.8e-3 with the default fp16 block-exponent gradient chain (the plain-bf16
    chain of TDET_INTERNAL_DTYPE=bf16 random-walks to 1.1-2.6e-2 at this depth: one 2^-9 rounding per
    stored gradient tensor and per scale-folded dgrad operand, tools/diag_grads.py);
  * arithmetic check of the neck (its gradient chain is plain bf16 and short): the 16 FPN gradients
    vs the oracle with rounding hooks where the kernels round (kernel_rounding=True): <= 1.5e-3
    (measured 2-4e-4: fp32 accumulation order + the few bf16 near-ties it flips);
  * the 16 FPN gradients also vs the plain fp32 oracle (<= 2e-2; measured ~5e-3);
  * local consistency of the training forward: each P level vs an fp32 recomputation from the
    kernels' own stored laterals (<= 4e-3, one bf16 rounding);
  * reported, not gated: backbone gradients vs the plain fp32 oracle (ReLU mask flips of the 16-bit
    forward; 0.1-0.25, cosine >= 0.97 -- PyTorch's own bf16 autocast is worse) .
"""
import pytest
import torch

from oracle import grad_oracle, resnet_fpn_oracle as orc
from tests import helpers

pytestmark = pytest.mark.gpu

GATE = 1e-2


def _train_pair(depth, seed, dev, bnstats, frozen_stages=1, bn_frozen=True):
    bb, neck = helpers.build_product_pair(depth, seed=seed, bnstats=bnstats, frozen_stages=frozen_stages,
                                          bn_eval=True, bn_frozen=bn_frozen)
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    bb = bb.to(dev)
    neck = neck.to(dev)
    bb.train()
    neck.train()
    return bb, neck, bsd, nsd


def _backbone_weight_dtype():
    from torch_detection_b200.models.backbone import resnet
    return resnet.INTERNAL_DTYPE


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


@pytest.mark.parametrize("depth,shape,bnstats,frozen,bn_frozen", [
    (50, (2, 3, 128, 160), False, 1, True),
    (50, (2, 3, 96, 128), True, 1, True),
    (50, (1, 3, 128, 128), True, 0, True),
    (18, (2, 3, 128, 160), True, 1, True),
    (101, (1, 3, 64, 96), False, 2, True),
    (50, (2, 3, 128, 160), True, 1, False),   # the reference's default bn_frozen=False: BN affine gradients
    (18, (2, 3, 96, 128), True, 0, False),
    (50, (2, 3, 96, 128), True, -1, False),   # the reference's constructor defaults: everything trains, stem included
    (18, (1, 3, 128, 160), True, -1, True),
])
def test_gradients_match_oracle(cuda_device, depth, shape, bnstats, frozen, bn_frozen):
    dev = cuda_device
    bb, neck, bsd, nsd = _train_pair(depth, 21, dev, bnstats, frozen, bn_frozen)
    bn_affine = not bn_frozen
    g = torch.Generator().manual_seed(5)
    x = torch.randn(*shape, generator=g).to(torch.bfloat16)
    feats = bb(x.to(dev))
    outs = neck(feats)
    assert all(o.requires_grad for o in outs)
    grads = [torch.randn(o.shape, generator=g).to(torch.bfloat16) for o in outs]
    torch.autograd.backward(list(outs), [t.to(dev).contiguous(memory_format=torch.channels_last) for t in grads])
    torch.cuda.synchronize()
    got_b = {k: p.grad.detach().cpu() for k, p in bb.named_parameters() if p.grad is not None}
    got_n = {k: p.grad.detach().cpu() for k, p in neck.named_parameters() if p.grad is not None}

    saved_b, saved_n = bb.saved_activations(), neck.saved_activations()
    wdt = _backbone_weight_dtype()
    tb, tn, tf_feats, tf_outs = grad_oracle.teacher_forced_grads(bsd, nsd, saved_b, saved_n, depth, grads,
                                                                 train_from_stage=frozen, kernel_rounding=True,
                                                                 bb_weight_dtype=wdt, x=x.float())
    xb, xn, _, _ = grad_oracle.teacher_forced_grads(bsd, nsd, saved_b, saved_n, depth, grads,
                                                    train_from_stage=frozen, bb_weight_dtype=wdt, bn_affine=bn_affine,
                                                    x=x.float())
    pb, pn, _, _ = grad_oracle.plain_grads(bsd, nsd, x.float(), depth, grads, train_from_stage=frozen,
                                           bn_affine=bn_affine)
    assert set(got_b) == set(xb), (sorted(set(got_b) ^ set(xb))[:8])
    if bn_affine:
        bnk = [k for k in xb if ".bn" in k or ".downsample.1." in k]
        worst = max(bnk, key=lambda k: orc.rel_l2(got_b[k], xb[k]))
        print("BN affine grads (%d): max rel-L2 %.2e (%s)" % (len(bnk), orc.rel_l2(got_b[worst], xb[worst]), worst))
    assert set(got_n) == set(tn) and len(got_n) == 16
    # the forced forward reproduces the CUDA outputs up to ONE layer of arithmetic (fp32 conv of the
    # kernels' own stored inputs): every stored tensor is locally consistent with its inputs
    fwd = [orc.rel_l2(a.float(), b) for a, b in zip(outs, tf_outs)]
    print("train-mode P levels vs one-layer fp32 recomputation:", ["%.2e" % e for e in fwd])
    assert max(fwd) <= 4e-3
    errs_n = {k: orc.rel_l2(got_n[k], tn[k]) for k in tn}
    errs_b = {k: orc.rel_l2(got_b[k], tb[k]) for k in tb if k in got_b}
    plain_n = {k: orc.rel_l2(got_n[k], pn[k]) for k in pn}
    plain_b = {k: (orc.rel_l2(got_b[k], pb[k]), _cos(got_b[k], pb[k])) for k in pb}
    exact_n = {k: orc.rel_l2(got_n[k], xn[k]) for k in xn}
    exact_b = {k: orc.rel_l2(got_b[k], xb[k]) for k in xb}
    print("FPN grads vs teacher-forced (kernel rounding): max %.2e ; (exact backward): max %.2e ; vs plain "
          "fp32: max %.2e" % (max(errs_n.values()), max(exact_n.values()), max(plain_n.values())))
    print("backbone grads (%d) vs teacher-forced (kernel rounding): max %.2e median %.2e ; (exact backward): "
          "max %.2e median %.2e" %
          (len(errs_b), max(errs_b.values()), sorted(errs_b.values())[len(errs_b) // 2],
           max(exact_b.values()), sorted(exact_b.values())[len(exact_b) // 2]))
    print("backbone grads vs plain fp32 (report only): max rel-L2 %.2e, min cosine %.4f" %
          (max(v[0] for v in plain_b.values()), min(v[1] for v in plain_b.values())))
    bad = {k: v for k, v in errs_n.items() if not v <= 1.5e-3}
    assert not bad, "FPN gradients over 1.5e-3 vs the teacher-forced oracle with kernel rounding: %s" % bad
    bad = {k: v for k, v in list(exact_n.items()) + list(exact_b.items()) if not v <= GATE}
    assert not bad, "gradients over 1e-2 vs the exact teacher-forced backward: %s" % bad
    bad = {k: v for k, v in plain_n.items() if not v <= 2e-2}
    assert not bad, "FPN gradients over 2e-2 vs the plain fp32 oracle: %s" % bad
    # frozen parameters received nothing
    assert all(p.grad is None for k, p in bb.named_parameters() if not p.requires_grad)


def test_second_step_uses_updated_weights(cuda_device):
    """An optimizer step changes the parameters in place: derived operands are refreshed in place
    (plans survive) and the next step's gradients follow the new weights."""
    dev = cuda_device
    bb, neck, _, _ = _train_pair(50, 3, dev, True)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 3, 96, 128, generator=g).to(torch.bfloat16).to(dev)
    params = [p for p in list(bb.parameters()) + list(neck.parameters()) if p.requires_grad]
    opt = torch.optim.SGD(params, lr=1e-3)
    n_plans = None
    for step in range(3):
        opt.zero_grad(set_to_none=True)
        outs = neck(bb(x))
        loss = 1e-4 * sum((o.float() ** 2).mean() for o in outs)  # ~1% relative weight update per step
        loss.backward()
        opt.step()
        if step == 0:
            n_plans = (len(bb._plans), len(neck._plans))
    torch.cuda.synchronize()
    assert (len(bb._plans), len(neck._plans)) == n_plans, "plans were rebuilt after an optimizer step"
    # gradients of the final state against the oracle on the final weights
    opt.zero_grad(set_to_none=True)
    outs = neck(bb(x))
    grads = [torch.randn(o.shape, generator=g).to(torch.bfloat16) for o in outs]
    torch.autograd.backward(list(outs), [t.to(dev).contiguous(memory_format=torch.channels_last) for t in grads])
    torch.cuda.synchronize()
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    tb, tn, _, _ = grad_oracle.teacher_forced_grads(bsd, nsd, bb.saved_activations(),
                                                    neck.saved_activations(), 50, grads, kernel_rounding=True,
                                                    bb_weight_dtype=_backbone_weight_dtype())
    for k, p in neck.named_parameters():
        assert orc.rel_l2(p.grad.cpu(), tn[k]) <= 1.5e-3, k
    xb, _, _, _ = grad_oracle.teacher_forced_grads(bsd, nsd, bb.saved_activations(), neck.saved_activations(),
                                                   50, grads, bb_weight_dtype=_backbone_weight_dtype())
    for k, p in bb.named_parameters():
        if p.grad is not None:
            assert orc.rel_l2(p.grad.cpu(), xb[k]) <= GATE, k


def test_eval_mode_unchanged_and_unsupported_training_configs(cuda_device):
    dev = cuda_device
    bb, neck = helpers.build_product_pair(50, seed=0)
    bb = bb.to(dev)
    x = torch.randn(1, 3, 64, 64).to(dev)
    bb.eval()
    assert not bb(x)[0].requires_grad          # eval mode never builds a graph
    bb.train()                                   # frozen_stages=-1 (the reference default): the stem trains too
    assert bb(x)[0].requires_grad
    bb2, _ = helpers.build_product_pair(50, seed=0, frozen_stages=1, bn_eval=False)
    bb2 = bb2.to(dev).train()
    with pytest.raises(NotImplementedError):   # batch-statistics BatchNorm
        bb2(x)


def test_retinanet_style_neck_gradients(cuda_device):
    """FPN(start_level=1, add_extra_convs=True): extra stride-2 convs on C5 with the reference's in-place ReLU
    between them (fpn.py:118-124); all neck gradients and the backbone's vs the plain fp32 oracle's FPN part
    (no ReLU inside the neck except that one) and the mask-matched backbone check through C-level agreement."""
    from torch_detection_b200.models.necks import FPN
    dev = cuda_device
    bb, _ = helpers.build_product_pair(50, seed=9, bnstats=True, frozen_stages=1, bn_eval=True, bn_frozen=True)
    torch.manual_seed(9)
    neck = FPN([256, 512, 1024, 2048], 256, 5, start_level=1, add_extra_convs=True)
    neck.init_weights()
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    bb, neck = bb.to(dev).train(), neck.to(dev).train()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 128, 160, generator=g).to(torch.bfloat16)
    feats = bb(x.to(dev))
    outs = neck(feats)
    assert len(outs) == 5
    grads = [torch.randn(o.shape, generator=g).to(torch.bfloat16) for o in outs]
    torch.autograd.backward(list(outs), [t.to(dev).contiguous(memory_format=torch.channels_last) for t in grads])
    torch.cuda.synchronize()
    # neck-only oracle on the kernels' own C levels: fp32 autograd of the reference's FPN forward
    cs = [f.detach().float().cpu().requires_grad_(True) for f in feats]
    leaf = {k: v.clone().float().requires_grad_(True) for k, v in nsd.items()}
    ref_outs = list(orc.fpn_forward(leaf, [c for c in cs], [256, 512, 1024, 2048], 256, 3, start_level=1))
    # extra levels with the ReLU decision of the kernels' own returned P6 (mask-matched, as for the backbone)
    p6 = torch.nn.functional.conv2d(cs[3], leaf["fpn_convs.3.conv.weight"], leaf["fpn_convs.3.conv.bias"], 2, 1)
    p6 = grad_oracle._force(p6, outs[3].detach().float().cpu(), True)
    p7 = torch.nn.functional.conv2d(p6, leaf["fpn_convs.4.conv.weight"], leaf["fpn_convs.4.conv.bias"], 2, 1)
    ref_outs += [p6, p7]
    for a, b in zip(outs, ref_outs):
        assert orc.rel_l2(a.float(), b) <= GATE   # two stored bf16 tensors + bf16 weights deep
    torch.autograd.backward(ref_outs, [t.float() for t in grads])
    for k, p in neck.named_parameters():
        e = orc.rel_l2(p.grad.cpu(), leaf[k].grad)
        assert e <= GATE, (k, e)
    # the gradient handed to the backbone for C3..C5 (C2 is unused with start_level=1): check through layer4
    got = {k: p.grad.detach().cpu() for k, p in bb.named_parameters() if p.grad is not None}
    pb, _, _, _ = grad_oracle.plain_grads(bsd, nsd, x.float(), 50, grads, train_from_stage=1, start_level=1,
                                          add_extra_convs=True)
    assert set(got) == set(pb)
    cos = min(torch.nn.functional.cosine_similarity(got[k].flatten().double(), pb[k].flatten().double(), dim=0).item()
              for k in pb if k.startswith("layer4"))
    print("RetinaNet-style neck: layer4 gradient cosine vs plain fp32 oracle >= %.4f" % cos)
    assert cos >= 0.97
