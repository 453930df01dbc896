"""Gradient oracle for the training configuration (TEST INFRASTRUCTURE ONLY: imported by tests/,
never by the product path).

The reference trains the path with torch autograd over its nn.Conv2d / nn.BatchNorm2d(eval) /
ReLU / add modules (models/backbone/resnet.py:97-119,253-268; models/necks/fpn.py:88-125).  Two
restatements, both plain CPU fp32 autograd over the oracle's forward:

  plain_grads           autograd through oracle.resnet_fpn_oracle.resnet_fpn_forward as is (the
                        reference's arithmetic, pinned bit-exact to the reference's forward by
                        tests/test_oracle_vs_reference.py; its backward is ATen's own autograd
                        formulas of the same ops).
  teacher_forced_grads  the same graph with bf16 rounding placed where the CUDA training path rounds
                        (conv weights, the image, every stored activation), passed straight through
                        in backward.  ReLU masks then come from the bf16 forward -- the comparison
                        SURVEY.md 8(c) calls "teacher-forced": bf16 flips a small fraction of ReLU
                        decisions relative to fp32, which changes backbone gradients by O(0.1) in
                        rel-L2 for ANY bf16 implementation (PyTorch's own included), so the gate
                        for backbone gradients must match masks.

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


def trainable_backbone_key(train_from_stage):
    """Conv weights of layer{train_from_stage+1..4} (0-based stage index), BN frozen."""
    def wanted(k):
        if not k.startswith("layer") or not k.endswith(".weight"):
            return False
        if ".bn" in k or ".downsample.1." in k:
            return False
        return int(k[5]) - 1 >= train_from_stage
    return wanted


def plain_grads(bb_sd, neck_sd, x, depth, grad_outs, train_from_stage=1, out_channels=256, num_outs=5):
    bb = _leafify(bb_sd, trainable_backbone_key(train_from_stage))
    neck = _leafify(neck_sd, lambda k: True)
    feats, outs = orc.resnet_fpn_forward(bb, neck, x.float(), depth, out_channels, num_outs)
    torch.autograd.backward(list(outs), [g.float() for g in grad_outs])
    gb = {k: v.grad for k, v in bb.items() if v.requires_grad}
    gn = {k: v.grad for k, v in neck.items() if v.requires_grad}
    return gb, gn, feats, outs


def _r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _ste(t):
    """bf16 rounding in forward, identity in backward."""
    return t + (_r(t) - t).detach()


def teacher_forced_grads(bb_sd, neck_sd, x, depth, grad_outs, train_from_stage=1, out_channels=256,
                         num_outs=5):
    kind, counts = orc.ARCH[depth]
    bb = _leafify(bb_sd, trainable_backbone_key(train_from_stage))
    neck = _leafify(neck_sd, lambda k: True)

    def w(sd, k):
        return _ste(sd[k])

    def cbn(inp, wkey, bnp, stride=1, pad=0, relu=False, res=None):
        y = F.conv2d(inp, w(bb, wkey), None, stride, pad)
        scale = bb[bnp + ".weight"] / torch.sqrt(bb[bnp + ".running_var"] + orc.BN_EPS)
        shift = bb[bnp + ".bias"] - bb[bnp + ".running_mean"] * scale
        y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        if res is not None:
            y = y + res
        if relu:
            y = F.relu(y)
        return _ste(y)

    h = cbn(_r(x.float()), "conv1.weight", "bn1", 2, 3, relu=True)
    h = F.max_pool2d(h, 3, 2, 1)
    feats = []
    for li, nblocks in enumerate(counts):
        for b in range(nblocks):
            p = "layer%d.%d" % (li + 1, b)
            s = 2 if (b == 0 and li > 0) else 1
            res = h
            if (p + ".downsample.0.weight") in bb:
                res = cbn(h, p + ".downsample.0.weight", p + ".downsample.1", s, 0)
            if kind == "bottleneck":
                o = cbn(h, p + ".conv1.weight", p + ".bn1", 1, 0, relu=True)
                o = cbn(o, p + ".conv2.weight", p + ".bn2", s, 1, relu=True)
                h = cbn(o, p + ".conv3.weight", p + ".bn3", 1, 0, relu=True, res=res)
            else:
                o = cbn(h, p + ".conv1.weight", p + ".bn1", s, 1, relu=True)
                h = cbn(o, p + ".conv2.weight", p + ".bn2", 1, 1, relu=True, res=res)
        feats.append(h)
    n = len(feats)
    lats = [None] * n
    for j in range(n - 1, -1, -1):
        y = F.conv2d(feats[j], w(neck, "lateral_convs.%d.conv.weight" % j),
                     neck["lateral_convs.%d.conv.bias" % j])
        if j < n - 1:
            y = y + F.interpolate(lats[j + 1], scale_factor=2, mode="nearest")
        lats[j] = _ste(y)
    outs = [_ste(F.conv2d(lats[j], w(neck, "fpn_convs.%d.conv.weight" % j),
                          neck["fpn_convs.%d.conv.bias" % j], 1, 1)) for j in range(n)]
    for _ in range(num_outs - n):
        outs.append(F.max_pool2d(outs[-1], 1, stride=2))
    torch.autograd.backward(list(outs), [g.float() for g in grad_outs])
    gb = {k: v.grad for k, v in bb.items() if v.requires_grad}
    gn = {k: v.grad for k, v in neck.items() if v.requires_grad}
    return gb, gn, feats, outs
