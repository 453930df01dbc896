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
    # end to end against the PLAIN fp32 reference autograd (different ReLU masks, so rel-L2 is ill-posed: SURVEY F11):
    # the direction of every backbone gradient must still agree -- a systematic forward/backward mismatch that
    # teacher forcing would hide fails here
    low = {k: v[1] for k, v in plain_b.items() if not v[1] >= 0.95}
    assert not low, "backbone gradients whose cosine vs the plain fp32 oracle is below 0.95: %s" % low
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


def _backward_once(bb, neck, x, grads, dev):
    outs = neck(bb(x.to(dev)))
    torch.autograd.backward(list(outs), [t.to(dev).contiguous(memory_format=torch.channels_last) for t in grads])
    torch.cuda.synchronize()
    return outs


def test_train_full_size(cuda_device):
    """BASELINE.json config 4 at its real geometry (2 images of 800x1333 zero-padded to 800x1344, frozen BN, stem
    and stage 1 frozen): all 58 parameter gradients (42 backbone conv weights + 16 FPN) within 1e-2 of the exact fp32
    backward over the kernels' own stored activations (mask-matched).  This is the size at which the
    wgrad cost model picks one-wave split-K and the 128 x 128 tiles, the stride-2 dgrads run as parity classes over
    100 x 168 / 50 x 84 / 25 x 42 gradients and the long-K dgrads run as CTA pairs -- none of which the small
    cases reach."""
    dev = cuda_device
    bb, neck, bsd, nsd = _train_pair(50, 21, dev, True, 1, True)
    g = torch.Generator().manual_seed(5)
    x = torch.zeros(2, 3, 800, 1344)
    x[:, :, :, :1333] = torch.randn(2, 3, 800, 1333, generator=g)
    x = x.to(torch.bfloat16)
    with torch.no_grad():
        shapes = [(2, 256, 200 >> i, 336 >> i) for i in range(4)] + [(2, 256, 13, 21)]
    grads = [torch.randn(s, generator=g).to(torch.bfloat16) for s in shapes]
    outs = _backward_once(bb, neck, x, grads, dev)
    assert [tuple(o.shape) for o in outs] == shapes
    got_b = {k: p.grad.detach().cpu() for k, p in bb.named_parameters() if p.grad is not None}
    got_n = {k: p.grad.detach().cpu() for k, p in neck.named_parameters() if p.grad is not None}
    assert len(got_b) == 42 and len(got_n) == 16
    # the plans really are the full-size ones: parity-class stride-2 dgrads, small-tile and large-tile wgrads, CTA pairs
    info = bb._last_bwd_run[0].launch_info()
    kinds = [l["kind"] for l in info]
    assert kinds.count(17) == 3, "three stride-2 3x3 dgrads as parity classes + merge"
    wg = [(l["tile_n"], l["variant"]) for l in info + neck._last_bwd_run[0].launch_info() if l["kind"] == 5]
    assert any(v == 1281 for _, v in wg) and any(t == 256 for t, _ in wg), wg   # 128 x 128 and 256-wide wgrad tiles
    assert any(l["kind"] == 3 and (l["variant"] & 8192) for l in info), "no CTA-pair dgrad in the full-size plan"
    saved_b, saved_n = bb.saved_activations(), neck.saved_activations()
    xb, xn, _, tf_outs = grad_oracle.teacher_forced_grads(bsd, nsd, saved_b, saved_n, 50, grads, train_from_stage=1,
                                                          bb_weight_dtype=_backbone_weight_dtype(), x=x.float())
    del saved_b, saved_n
    fwd = [orc.rel_l2(a.float(), b) for a, b in zip(outs, tf_outs)]
    assert max(fwd) <= 4e-3, fwd
    errs = {k: orc.rel_l2(got_b[k], xb[k]) for k in xb}
    errs.update({"neck." + k: orc.rel_l2(got_n[k], xn[k]) for k in xn})
    worst = max(errs, key=errs.get)
    print("full-size gradients (%d): max rel-L2 %.2e (%s), median %.2e" %
          (len(errs), errs[worst], worst, sorted(errs.values())[len(errs) // 2]))
    bad = {k: v for k, v in errs.items() if not v <= GATE}
    assert len(errs) == 58 and not bad, bad


def test_training_plans_under_an_sm_reserve(cuda_device):
    """The multi-GPU training plans are built with 8 SMs left to NCCL (BucketAllReduce.attach): 140-CTA grids and
    even-rounded CTA-pair grids give the same gradients as full-width ones (fp32 accumulation order aside)."""
    from torch_detection_b200 import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 256, 320, generator=g).to(torch.bfloat16)
    res = []
    for reserve in (0, 8):
        bb, neck, bsd, nsd = _train_pair(50, 21, dev, True, 1, True)
        try:
            engine.set_sm_reserve(dev, reserve)
            outs = neck(bb(x.to(dev)))
            if not res:
                grads = [torch.randn(o.shape, generator=g).to(torch.bfloat16) for o in outs]
            torch.autograd.backward(list(outs), [t.to(dev).contiguous(memory_format=torch.channels_last) for t in grads])
            torch.cuda.synchronize()
            # (persistent conv / dgrad grids; the wgrad grids are tiles x split-K slices sized for <= 2 waves of the
            # reduced SM count)
            grids = [l["grid"] for l in bb._last_bwd_run[0].launch_info() + bb._last_run[0].launch_info()
                     if l["kind"] in (1, 3)]
        finally:
            engine.set_sm_reserve(dev, 0)
        res.append(({k: p.grad.detach().cpu() for k, p in list(bb.named_parameters()) + list(neck.named_parameters())
                     if p.grad is not None}, [o.detach().cpu() for o in outs], grids))
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    assert max(res[1][2]) <= sms - 8 < max(res[0][2]), (max(res[0][2]), max(res[1][2]))
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.equal(a, b)
    worst = max(orc.rel_l2(res[1][0][k], res[0][0][k]) for k in res[0][0])
    print("gradients under an 8-SM reserve vs full-width grids: max rel-L2 %.2e" % worst)
    assert worst <= 1e-4
    # and against the oracle
    xb, xn, _, _ = grad_oracle.teacher_forced_grads(bsd, nsd, bb.saved_activations(), neck.saved_activations(), 50,
                                                    grads, train_from_stage=1, bb_weight_dtype=_backbone_weight_dtype())
    for k, v in xb.items():
        assert orc.rel_l2(res[1][0][k], v) <= GATE, k


def test_trainable_neck_on_a_frozen_fp32_backbone(cuda_device):
    """A frozen / eval backbone fed fp32 images runs the split-precision path and hands over fp32 features that
    carry their hi|lo pairs; a TRAINABLE neck must not follow them into the split path (its backward plan needs the
    bf16 laterals): forward, backward and gradients against fp32 autograd of the reference's FPN on the same inputs."""
    dev = cuda_device
    bb, neck = helpers.build_product_pair(50, seed=14, bnstats=True)
    nsd = helpers.cpu_state(neck)
    bb = bb.to(dev).eval()
    neck = neck.to(dev).train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 128, 160, generator=g)
    with torch.no_grad():
        feats = bb(x.to(dev))
    assert all(f.dtype == torch.float32 and getattr(f, "_tdet_split", None) is not None for f in feats)
    for rep in range(2):   # the second call used to reuse a stale feature list
        for p in neck.parameters():
            p.grad = None
        outs = neck(feats)
        grads = [torch.randn(o.shape, generator=g).to(torch.bfloat16) for o in outs]
        torch.autograd.backward(list(outs), [t.to(dev).contiguous(memory_format=torch.channels_last) for t in grads])
        torch.cuda.synchronize()
    cs = [f.detach().to(torch.bfloat16).float().cpu() for f in feats]
    leaf = {k: v.clone().float().requires_grad_(True) for k, v in nsd.items()}
    ref = orc.fpn_forward(leaf, cs, [256, 512, 1024, 2048], 256, 5)
    for a, b in zip(outs, ref):
        assert orc.rel_l2(a.float(), b.detach()) <= GATE
    torch.autograd.backward(list(ref), [t.float() for t in grads])
    for k, p in neck.named_parameters():
        e = orc.rel_l2(p.grad.cpu(), leaf[k].grad)
        assert e <= 2e-2, (k, e)


def test_data_updates_need_invalidate_operands(cuda_device):
    """Parameter changes are detected through tensor._version; an in-place update through .data does not bump it
    (legacy optimizers, EMA): invalidate_operands() re-derives the packed weights."""
    dev = cuda_device
    bb, neck = helpers.build_product_pair(18, seed=0)
    bb, neck = bb.to(dev).eval(), neck.to(dev).eval()
    x = torch.randn(1, 3, 64, 64).to(torch.bfloat16).to(dev)
    with torch.no_grad():
        f0 = bb(x)
        v0 = bb.layer1[0].conv1.weight._version
        bb.layer1[0].conv1.weight.data.mul_(0.5)
        neck.lateral_convs[0].conv.weight.data.mul_(0.5)
        assert bb.layer1[0].conv1.weight._version == v0
        bb.invalidate_operands()
        neck.invalidate_operands()
        f1 = bb(x)
        p1 = neck(f1)
    want_f, want_p = orc.resnet_fpn_forward(helpers.cpu_state(bb), helpers.cpu_state(neck), x.float().cpu(), 18)
    assert not torch.equal(f0[0], f1[0])
    for a, b in zip(list(f1) + list(p1), list(want_f) + list(want_p)):
        assert orc.rel_l2(a.float(), b) <= GATE


def test_gradient_accumulation_with_deferred_buckets(cuda_device):
    """BucketAllReduce(defer=True) hands side-stream copies to autograd; when the parameters already hold a gradient
    autograd ACCUMULATES into it on the compute stream, which must be ordered after the copy / collective: two
    backward passes without zero_grad give exactly twice the gradient."""
    from torch_detection_b200 import training
    dev = cuda_device
    bb, neck, _, _ = _train_pair(50, 3, dev, True)
    sync = training.BucketAllReduce(defer=True)
    bb.set_grad_sync(sync)
    neck.set_grad_sync(sync)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 3, 96, 128, generator=g).to(torch.bfloat16)
    outs = neck(bb(x.to(dev)))
    grads = [torch.randn(o.shape, generator=g).to(torch.bfloat16) for o in outs]
    params = [p for p in list(bb.parameters()) + list(neck.parameters()) if p.requires_grad]
    torch.autograd.backward(list(outs), [t.to(dev).contiguous(memory_format=torch.channels_last) for t in grads])
    sync.finish()
    torch.cuda.synchronize()
    once = [p.grad.detach().clone() for p in params]
    assert sync.buckets_not_overlapped == 0
    _backward_once(bb, neck, x, grads, dev)     # no zero_grad: accumulates
    sync.finish()
    torch.cuda.synchronize()
    assert sync.buckets_not_overlapped > 0
    for p, a in zip(params, once):
        assert orc.rel_l2(p.grad, 2 * a) <= 1e-5


@pytest.mark.parametrize("activation", [None, "relu"])
def test_pafpn_gradients(cuda_device, activation):
    """SURVEY 8(f) row f3, training: PAFPN's bottom-up path (pafpn.py:131-134) backward -- every neck parameter
    gradient and the gradients handed back for C2..C5 against fp32 autograd of the oracle's PAFPN forward on the
    same bf16 inputs.  With activation='relu' the comparison is mask-matched like the backbone's: the
    oracle's ReLUs take the decisions (and forward values) of the kernels' own stored t_j / N_j, so that a handful of
    flipped elements does not hide -- or fake -- a backward error; the plain comparison's cosine is asserted too."""
    from torch_detection_b200 import models
    from torch_detection_b200.utils import obj_from_dict
    dev = cuda_device
    torch.manual_seed(21)
    neck = obj_from_dict(dict(type="PAFPN", in_channels=[256, 512, 1024, 2048], out_channels=256, num_outs=5,
                              activation=activation), parent=models.necks)
    neck.init_weights()
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():   # biases away from zero
        for p in neck.parameters():
            if p.dim() == 1:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    nsd = helpers.cpu_state(neck)
    neck = neck.to(dev).train()
    shapes = [(2, 256, 40, 56), (2, 512, 20, 28), (2, 1024, 10, 14), (2, 2048, 5, 7)]
    cs = [torch.randn(s, generator=g).abs().to(torch.bfloat16) for s in shapes]
    feats = [c.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True) for c in cs]
    for rep in range(2):
        for p in neck.parameters():
            p.grad = None
        for f in feats:
            f.grad = None
        outs = neck(feats)
        assert len(outs) == 5
        grads = [torch.randn(o.shape, generator=torch.Generator().manual_seed(9 + i)).to(torch.bfloat16)
                 for i, o in enumerate(outs)]
        torch.autograd.backward(list(outs), [t.to(dev).contiguous(memory_format=torch.channels_last) for t in grads])
        torch.cuda.synchronize()
    leaf = {k: v.clone().float().requires_grad_(True) for k, v in nsd.items()}
    cl = [c.float().requires_grad_(True) for c in cs]
    ref = orc.pafpn_forward(leaf, cl, [256, 512, 1024, 2048], 256, 5, activation=activation)
    for a, b in zip(outs, ref):
        assert orc.rel_l2(a.float(), b.detach()) <= GATE
    if activation == "relu":
        plain = torch.autograd.grad(list(ref), [leaf[k] for k, _ in neck.named_parameters()], [t.float() for t in grads])
        for (k, p), gr in zip(neck.named_parameters(), plain):
            assert _cos(p.grad.cpu(), gr) >= 0.99, (k, _cos(p.grad.cpu(), gr))
        # mask-matched oracle: the same forward with the kernels' own ReLU decisions
        saved = neck._train_state["plan"].pa

        def stored(act):
            n, h, w, c = act.shape
            return act.buf[act.offset:act.offset + n * h * w * c].view(n, h, w, c).permute(0, 3, 1, 2).float().cpu()

        leaf = {k: v.clone().float().requires_grad_(True) for k, v in nsd.items()}
        cl = [c.float().requires_grad_(True) for c in cs]
        pyr = list(orc.fpn_forward(leaf, cl, [256, 512, 1024, 2048], 256, 4))
        conv = torch.nn.functional.conv2d
        ref = [pyr[0]]
        for j in range(1, 4):
            t = conv(ref[-1], leaf["pa_convs1.%d.conv.weight" % (j - 1)], leaf["pa_convs1.%d.conv.bias" % (j - 1)], 2, 1)
            t = grad_oracle._force(t, stored(saved["t"][j]), True)
            nj = conv(pyr[j] + t, leaf["pa_convs2.%d.conv.weight" % (j - 1)], leaf["pa_convs2.%d.conv.bias" % (j - 1)], 1, 1)
            ref.append(grad_oracle._force(nj, outs[j].detach().float().cpu(), True))
        ref.append(torch.nn.functional.max_pool2d(ref[-1], 1, stride=2))
    torch.autograd.backward(list(ref), [t.float() for t in grads])
    worst = 0.0
    tol = 1e-2
    for k, p in neck.named_parameters():
        assert p.grad is not None, k
        e = orc.rel_l2(p.grad.cpu(), leaf[k].grad)
        worst = max(worst, e)
        assert e <= tol, (k, e)
    for j, (f, c) in enumerate(zip(feats, cl)):
        e = orc.rel_l2(f.grad.float().cpu(), c.grad)
        worst = max(worst, e)
        assert e <= tol, ("C%d" % (j + 2), e)
    print("PAFPN %s gradients: worst rel-L2 %.2e" % (activation, worst))
    if activation == "relu":
        neck6 = obj_from_dict(dict(type="PAFPN", in_channels=[256, 512, 1024, 2048], out_channels=256, num_outs=5,
                                   activation="relu6"), parent=models.necks).to(dev).train()
        with pytest.raises(NotImplementedError):
            neck6(feats)


@pytest.mark.parametrize("base_width,cardinality,frozen", [(4, 32, 1), (8, 16, 2)])
def test_resnext_gradients(cuda_device, base_width, cardinality, frozen):
    """SURVEY 8(f) row f4, training: ResNeXt-50 (grouped 3x3, models/backbone/resnext.py:84-87).  The data gradient of
    the grouped conv runs as a grouped conv over the block-diagonal operand, its weight gradient is accumulated dense
    and unpacked to the block diagonal; every gradient against the teacher-forced fp32 oracle (exact backward over the
    kernels' own stored activations, <= 1e-2) and, for direction, the plain fp32 oracle (cosine >= 0.95)."""
    from torch_detection_b200 import models
    from torch_detection_b200.utils import obj_from_dict
    dev = cuda_device
    torch.manual_seed(13)
    bb = obj_from_dict(dict(type="ResNeXt", depth=50, base_width=base_width, cardinality=cardinality,
                            frozen_stages=frozen, bn_eval=True, bn_frozen=True), parent=models.backbone)
    bb.init_weights()
    sd = bb.state_dict()
    orc.randomize_bn_stats(sd, generator=torch.Generator().manual_seed(3))
    bb.load_state_dict(sd)
    _, neck = helpers.build_product_pair(50, seed=5)
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    bb, neck = bb.to(dev).train(), neck.to(dev).train()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 96, 128, generator=g).to(torch.bfloat16)
    outs = neck(bb(x.to(dev)))
    grads = [torch.randn(o.shape, generator=g).to(torch.bfloat16) for o in outs]
    torch.autograd.backward(list(outs), [t.to(dev).contiguous(memory_format=torch.channels_last) for t in grads])
    torch.cuda.synchronize()
    got_b = {k: p.grad.detach().cpu() for k, p in bb.named_parameters() if p.grad is not None}
    assert all(got_b[k].shape == bsd[k].shape for k in got_b)
    grouped = [k for k in got_b if k.endswith("conv2.weight")]
    assert grouped and all(bsd[k].shape[1] * cardinality == bsd[k].shape[0] for k in grouped)
    xb, xn, _, _ = grad_oracle.teacher_forced_grads(bsd, nsd, bb.saved_activations(), neck.saved_activations(), 50, grads,
                                                    train_from_stage=frozen, bb_weight_dtype=_backbone_weight_dtype(),
                                                    x=x.float())
    pb, _, _, _ = grad_oracle.plain_grads(bsd, nsd, x.float(), 50, grads, train_from_stage=frozen)
    assert set(got_b) == set(xb)
    errs = {k: orc.rel_l2(got_b[k], xb[k]) for k in xb}
    worst = max(errs, key=errs.get)
    print("ResNeXt-50 %dx%dd gradients (%d): max rel-L2 %.2e (%s), grouped convs max %.2e" %
          (cardinality, base_width, len(errs), errs[worst], worst, max(errs[k] for k in grouped)))
    bad = {k: v for k, v in errs.items() if not v <= GATE}
    assert not bad, bad
    low = {k: _cos(got_b[k], pb[k]) for k in pb if not _cos(got_b[k], pb[k]) >= 0.95}
    assert not low, low
