"""Gradient oracle for the training configuration (TEST INFRASTRUCTURE ONLY: imported by tests/,
never by the product path).

The reference trains the path with torch autograd over its nn.Conv2d / nn.BatchNorm2d(eval) /
ReLU / add modules (models/backbone/resnet.py:97-119,253-268; models/necks/fpn.py:88-125).  Two
restatements, both plain CPU fp32 autograd over the oracle's forward:

  plain_grads           autograd through oracle.resnet_fpn_oracle.resnet_fpn_forward as is (the
                        reference's arithmetic, pinned bit-exact to the reference's forward by
                        tests/test_oracle_vs_reference.py; its backward is ATen's own autograd
                        formulas of the same ops).
  teacher_forced_grads  the same graph with every stored activation forced to the value the CUDA
                        forward stored (downloaded through the modules' saved_activations() test
                        API) and bf16-rounded conv weights, passed straight through in backward.
                        ReLU masks and wgrad inputs are then the kernels' own -- the comparison
                        SURVEY.md 8(c) calls "teacher-forced" / "mask-matched": any bf16 forward
                        flips ~1% of the ReLU decisions of the fp32 forward (and of any OTHER bf16
                        forward with a different accumulation order), which moves backbone gradients
                        by O(0.1) in rel-L2 -- measured here for an independent bf16 emulation:
                        0.07-0.18 -- for ANY bf16 implementation, PyTorch's own included.

Gradients are returned for every conv weight of the trainable stages (frozen BN: no affine
gradients; frozen stem / leading stages as in the reference's configs) and all FPN parameters.
"""
import torch
import torch.nn.functional as F

from . import resnet_fpn_oracle as orc


def _leafify(sd, wanted):
    out = {}
    for k, v in sd.items():
        t = v.detach().clone().float() if v.is_floating_point() else v.detach().clone()
        if wanted(k):
            t.requires_grad_(True)
        out[k] = t
    return out


def trainable_backbone_key(train_from_stage, bn_affine=False):
    """Conv weights of layer{train_from_stage+1..4} (0-based stage index); with bn_affine also the
    weight / bias of their (eval-mode) BatchNorms (the reference's bn_frozen=False)."""
    def wanted(k):
        if train_from_stage < 0 and (k == "conv1.weight" or (bn_affine and k in ("bn1.weight", "bn1.bias"))):
            return True   # the stem trains too (the reference's frozen_stages = -1)
        if not k.startswith("layer") or int(k[5]) - 1 < max(train_from_stage, 0):
            return False
        is_bn = ".bn" in k or ".downsample.1." in k
        if is_bn:
            return bn_affine and (k.endswith(".weight") or k.endswith(".bias"))
        return k.endswith(".weight")
    return wanted


def plain_grads(bb_sd, neck_sd, x, depth, grad_outs, train_from_stage=1, out_channels=256, num_outs=5,
                bn_affine=False, start_level=0, add_extra_convs=False):
    bb = _leafify(bb_sd, trainable_backbone_key(train_from_stage, bn_affine))
    neck = _leafify(neck_sd, lambda k: True)
    kind, _ = orc.ARCH[depth]
    in_ch = [64 * 2 ** i * orc.EXPANSION[kind] for i in range(4)]
    feats = orc.resnet_forward(bb, x.float(), depth)
    outs = orc.fpn_forward(neck, feats, in_ch, out_channels, num_outs, start_level=start_level,
                           add_extra_convs=add_extra_convs)
    torch.autograd.backward(list(outs), [g.float() for g in grad_outs])
    gb = {k: v.grad for k, v in bb.items() if v.requires_grad}
    gn = {k: v.grad for k, v in neck.items() if v.requires_grad}
    return gb, gn, feats, outs


def _r(t, dtype=torch.bfloat16):
    return t.to(dtype).to(torch.float32)


def _ste(t, dtype=torch.bfloat16):
    """16-bit rounding in forward, identity in backward."""
    return t + (_r(t, dtype) - t).detach()


def _force(y, stored, relu, round_grad=False):
    """Forward value := the CUDA path's stored activation, ReLU decision := its sign pattern;
    backward: the gradient of y under that decision.  round_grad: the total gradient w.r.t. the
    stored tensor passes through one bf16 rounding, as the kernel that materialises it does."""
    stored = stored.float()
    if relu:
        y = y * (stored > 0).to(y.dtype)
    out = y + (stored - y).detach()
    if round_grad and out.requires_grad:
        out.register_hook(_r)
    return out


def _tap(t, round_grad):
    """Identity whose gradient contribution is rounded to bf16 (a branch gradient the kernels store
    before merging it: shortcut dgrad, 2x2 sum-pooled lateral gradient, dC_k handed to the backbone)."""
    if not (round_grad and t.requires_grad):
        return t
    out = t * 1.0
    out.register_hook(_r)
    return out


class _ConvBNKernelModel(torch.autograd.Function):
    """conv + frozen BN with the CUDA path's operand rounding: forward uses bf16(w) and the fp32
    scale/shift epilogue; the data gradient uses the BN-scale-folded operand bf16(scale * w)
    (tdet_pack_dgrad_weight); the weight gradient is scale * (g (*) x) in fp32."""

    @staticmethod
    def forward(ctx, x, w, scale, shift, stride, pad, wdtype):
        ctx.save_for_backward(x, w, scale)
        ctx.conf = (stride, pad, wdtype)
        y = F.conv2d(x, _r(w, wdtype), None, stride, pad)
        return y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)

    @staticmethod
    def backward(ctx, g):
        x, w, scale = ctx.saved_tensors
        stride, pad, wdtype = ctx.conf
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = torch.nn.grad.conv2d_input(x.shape, _r(w * scale.view(-1, 1, 1, 1), wdtype), g, stride, pad)
        if ctx.needs_input_grad[1]:
            gw = torch.nn.grad.conv2d_weight(x, w.shape, g * scale.view(1, -1, 1, 1), stride, pad)
        return gx, gw, None, None, None, None, None


def teacher_forced_grads(bb_sd, neck_sd, saved_bb, saved_neck, depth, grad_outs, train_from_stage=1,
                         out_channels=256, num_outs=5, kernel_rounding=False, bb_weight_dtype=torch.bfloat16,
                         bn_affine=False, x=None):
    """fp32 autograd over the reference's graph with every stored activation (and hence every ReLU
    mask and every conv / wgrad input) forced to the value the CUDA training forward stored:
    `saved_bb` = ResNet.saved_activations(), `saved_neck` = FPN.saved_activations().  Conv weights are
    rounded to the kernels' operand format (straight-through): `bb_weight_dtype` for the backbone (fp16 with
    the default block-exponent activations, bf16 with TDET_INTERNAL_DTYPE=bf16), bf16 for the neck.

    kernel_rounding=False: the backward is exact fp32 -- what remains different from the CUDA
        backward is its bf16 storage of gradient tensors and dgrad operands (accumulates as
        ~sqrt(depth) * 2^-9) plus fp32 accumulation order.
    kernel_rounding=True: additionally one bf16 rounding at every point where the CUDA backward
        stores a gradient tensor, and the scale-folded dgrad operand: "rounding hooks placed exactly
        where the kernels round" (SURVEY.md 8c-2); what remains is fp32 accumulation order only, so
        this is the tight gate on the kernels' arithmetic."""
    kind, counts = orc.ARCH[depth]
    kr = kernel_rounding
    assert not (kr and bn_affine), "the kernel-rounding model covers conv weights only"
    bb = _leafify(bb_sd, trainable_backbone_key(train_from_stage, bn_affine))
    neck = _leafify(neck_sd, lambda k: True)

    def cbn(inp, wkey, bnp, stored, stride=1, pad=0, relu=False, res=None, round_grad=False, wdtype=None):
        wdtype = wdtype or bb_weight_dtype
        scale = bb[bnp + ".weight"] / torch.sqrt(bb[bnp + ".running_var"] + orc.BN_EPS)
        shift = bb[bnp + ".bias"] - bb[bnp + ".running_mean"] * scale
        groups = inp.shape[1] // bb[wkey].shape[1]   # > 1: ResNeXt's grouped 3x3 (models/backbone/resnext.py:84-87)
        if kr:
            assert groups == 1, "the kernel-rounding model covers dense convs"
            y = _ConvBNKernelModel.apply(inp, bb[wkey], scale, shift, stride, pad, wdtype)
        else:
            y = F.conv2d(inp, _ste(bb[wkey], wdtype), None, stride, pad, 1, groups)
            y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        if res is not None:
            y = y + res
        return y if stored is None else _force(y, stored, relu, round_grad)

    feats = []
    h = None
    if train_from_stage < 0:
        # trainable stem (frozen_stages = -1): bf16 image and bf16 stem weights, stem output forced to the stored
        # tensor; the max-pool then sees exactly the values the CUDA max-pool saw (same maxima, same tie rule)
        s_out = cbn(_r(x.float()), "conv1.weight", "bn1", saved_bb["stem.out"], 2, 3, relu=True, round_grad=kr,
                    wdtype=torch.bfloat16)
        h = F.max_pool2d(s_out, 3, 2, 1)
    for li, nblocks in enumerate(counts):
        if li < train_from_stage:
            feats.append(None)
            continue
        for b in range(nblocks):
            p = "layer%d.%d" % (li + 1, b)
            s = 2 if (b == 0 and li > 0) else 1
            if h is None:
                h = saved_bb[p + ".in"].float()  # frozen producer: a constant
            res = h
            if (p + ".downsample.0.weight") in bb:
                # the shortcut branch is stored in bf16 and consumed by conv3's epilogue; its data
                # gradient is stored (bf16) before it is merged into the block-input gradient
                res = _ste(cbn(_tap(h, kr), p + ".downsample.0.weight", p + ".downsample.1", None, s, 0))
            if kind == "bottleneck":
                o = cbn(h, p + ".conv1.weight", p + ".bn1", saved_bb[p + ".conv1"], 1, 0, relu=True, round_grad=kr)
                o = cbn(o, p + ".conv2.weight", p + ".bn2", saved_bb[p + ".conv2"], s, 1, relu=True, round_grad=kr)
                h = cbn(o, p + ".conv3.weight", p + ".bn3", saved_bb[p + ".conv3"], 1, 0, relu=True, res=res,
                        round_grad=kr)
            else:
                o = cbn(h, p + ".conv1.weight", p + ".bn1", saved_bb[p + ".conv1"], s, 1, relu=True, round_grad=kr)
                h = cbn(o, p + ".conv2.weight", p + ".bn2", saved_bb[p + ".conv2"], 1, 1, relu=True, res=res,
                        round_grad=kr)
        feats.append(h)
    n = len(feats)
    for j in range(n):
        if feats[j] is None:
            feats[j] = saved_neck["C%d" % j].float()  # frozen stage output: a constant for the neck
    lats = [None] * n
    for j in range(n - 1, -1, -1):
        # the neck consumes the returned bf16 feature map (a converted copy of the backbone's internal
        # stage output): forward value = what the neck stored, gradient straight through to the backbone
        cj = _force(feats[j], saved_neck["C%d" % j], False) if feats[j].requires_grad else feats[j]
        y = F.conv2d(_tap(cj, kr), _ste(neck["lateral_convs.%d.conv.weight" % j]),
                     neck["lateral_convs.%d.conv.bias" % j])
        if j < n - 1:
            y = y + F.interpolate(_tap(lats[j + 1], kr), scale_factor=2, mode="nearest")
        lats[j] = _force(y, saved_neck["lat%d" % j], False, kr)
    outs = [F.conv2d(lats[j], _ste(neck["fpn_convs.%d.conv.weight" % j]),
                     neck["fpn_convs.%d.conv.bias" % j], 1, 1) for j in range(n)]
    if kr and num_outs > n and outs[-1].requires_grad:
        outs[-1].register_hook(_r)  # dP of the coarsest level + scattered extra-level gradients: stored once
    for _ in range(num_outs - n):
        outs.append(F.max_pool2d(outs[-1], 1, stride=2))
    torch.autograd.backward(list(outs), [g.float() for g in grad_outs])
    gb = {k: v.grad for k, v in bb.items() if v.requires_grad}
    gn = {k: v.grad for k, v in neck.items() if v.requires_grad}
    return gb, gn, feats, outs


def oracle_saved_activations(bb_sd, neck_sd, x, depth):
    """The activations the CUDA training forward stores, computed by the fp32 oracle itself, under the
    names of ResNet.saved_activations() / FPN.saved_activations().  Forcing the oracle with its own
    activations must reproduce plain_grads (tests/test_grad_oracle_cpu.py)."""
    kind, counts = orc.ARCH[depth]
    saved_bb, saved_neck = {}, {}
    with torch.no_grad():
        h = F.conv2d(x.float(), bb_sd["conv1.weight"], None, 2, 3)
        h = F.relu(orc._bn(bb_sd, "bn1", h))
        h = F.max_pool2d(h, 3, 2, 1)
        feats = []
        for li, nblocks in enumerate(counts):
            for b in range(nblocks):
                p = "layer%d.%d" % (li + 1, b)
                s = 2 if (b == 0 and li > 0) else 1
                saved_bb[p + ".in"] = h
                res = h
                if (p + ".downsample.0.weight") in bb_sd:
                    res = orc._bn(bb_sd, p + ".downsample.1",
                                  F.conv2d(h, bb_sd[p + ".downsample.0.weight"], None, s))
                if kind == "bottleneck":
                    o = F.relu(orc._bn(bb_sd, p + ".bn1", F.conv2d(h, bb_sd[p + ".conv1.weight"])))
                    saved_bb[p + ".conv1"] = o
                    o = F.relu(orc._bn(bb_sd, p + ".bn2", F.conv2d(o, bb_sd[p + ".conv2.weight"], None, s, 1)))
                    saved_bb[p + ".conv2"] = o
                    h = F.relu(orc._bn(bb_sd, p + ".bn3", F.conv2d(o, bb_sd[p + ".conv3.weight"])) + res)
                    saved_bb[p + ".conv3"] = h
                else:
                    o = F.relu(orc._bn(bb_sd, p + ".bn1", F.conv2d(h, bb_sd[p + ".conv1.weight"], None, s, 1)))
                    saved_bb[p + ".conv1"] = o
                    h = F.relu(orc._bn(bb_sd, p + ".bn2", F.conv2d(o, bb_sd[p + ".conv2.weight"], None, 1, 1)) + res)
                    saved_bb[p + ".conv2"] = h
            feats.append(h)
        n = len(feats)
        lats = [None] * n
        for j in range(n - 1, -1, -1):
            y = F.conv2d(feats[j], neck_sd["lateral_convs.%d.conv.weight" % j],
                         neck_sd["lateral_convs.%d.conv.bias" % j])
            if j < n - 1:
                y = y + F.interpolate(lats[j + 1], scale_factor=2, mode="nearest")
            lats[j] = y
            saved_neck["C%d" % j] = feats[j]
            saved_neck["lat%d" % j] = y
    return saved_bb, saved_neck
