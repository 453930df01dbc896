"""CPU oracle for the ResNet + FPN feature-extraction path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* of the reference algorithm, used as the checker by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs.  Nothing under ``torch_detection_b200/`` may import it: the product path is the CUDA
extension and fails loudly without it.

Parity pinning: the reference ships no tests or golden vectors, so this
oracle is pinned by *executing the unmodified reference* (``/root/reference`` through the import shim
in ``oracle/reference_shim.py``) in the dev container:
  * ``tests/test_oracle_vs_reference.py`` asserts bit-identical outputs, state_dict keys/shapes and
    init statistics against the live reference whenever ``/root/reference`` exists;
  * ``oracle/make_golden.py`` ran the *reference itself* to write ``tests/golden/*.npz``; those
    fixtures travel to the GPU box where the reference does not exist, and
    ``tests/test_oracle_golden.py`` re-checks this oracle against them everywhere.

The restatement is functional (a ``state_dict`` in, tensors out) instead of a module tree, but it
issues the same ATen ops in the same order as the reference, which is what makes it bit-identical
on CPU:

  stem      conv 7x7/2 p3 -> BN(eval) -> ReLU -> maxpool 3x3/2 p1   models/backbone/resnet.py:253-258
  Bottleneck 1x1 -> BN -> ReLU -> 3x3(stride, pad=dil) -> BN -> ReLU -> 1x1 -> BN,
            (+ downsample 1x1 stride + BN), add, ReLU                models/backbone/resnet.py:97-119
  BasicBlock 3x3(stride) -> BN -> ReLU -> 3x3 -> BN, (+downsample), add, ReLU
                                                                     models/backbone/resnet.py:42-59
  stage loop / out_indices / bare tensor for one output              models/backbone/resnet.py:259-268
  FPN       laterals 1x1+b; top-down nearest x2 add (in place, coarse->fine); 3x3 p1 +b outputs;
            extra levels by stride-2 subsample or stride-2 convs      models/necks/fpn.py:88-125
  init      Kaiming-normal fan_out / BN gamma=1 beta=0; Xavier-uniform for FPN
                                   models/backbone/resnet.py:240-251, models/necks/fpn.py:80-86,
                                   models/utils/inits.py:5-46
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

# models/backbone/resnet.py:178-184
ARCH = {
    18: ("basic", (2, 2, 2, 2)),
    34: ("basic", (3, 4, 6, 3)),
    50: ("bottleneck", (3, 4, 6, 3)),
    101: ("bottleneck", (3, 4, 23, 3)),
    152: ("bottleneck", (3, 8, 36, 3)),
}
EXPANSION = {"basic": 1, "bottleneck": 4}
BN_EPS = 1e-5  # nn.BatchNorm2d default, models/utils/layers.py:50-54


# ----------------------------------------------------------------------------------------------
# parameter construction (reference init order matters: it consumes the global torch RNG in
# module-registration order, see make_resnet_state / make_fpn_state)
# ----------------------------------------------------------------------------------------------

def _bn_entries(sd, prefix, ch):
    sd[prefix + ".weight"] = torch.ones(ch)
    sd[prefix + ".bias"] = torch.zeros(ch)
    sd[prefix + ".running_mean"] = torch.zeros(ch)
    sd[prefix + ".running_var"] = torch.ones(ch)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def resnet_param_shapes(depth, num_stages=4, strides=(1, 2, 2, 2)):
    """Ordered (name, shape, kind) list == reference ``ResNet(depth).state_dict()`` order.

    kind is 'conv' or 'bn' (one entry per BN module; expands to 5 tensors)."""
    block, counts = ARCH[depth]
    exp = EXPANSION[block]
    out = [("conv1", (64, 3, 7, 7), "conv"), ("bn1", 64, "bn")]
    inplanes = 64
    for li, nblocks in enumerate(counts[:num_stages]):
        planes = 64 * 2 ** li
        stride = strides[li]
        for b in range(nblocks):
            p = "layer%d.%d" % (li + 1, b)
            cin = inplanes if b == 0 else planes * exp
            if block == "bottleneck":
                convs = [(planes, cin, 1, 1), (planes, planes, 3, 3), (planes * exp, planes, 1, 1)]
            else:
                convs = [(planes, cin, 3, 3), (planes, planes, 3, 3)]
            for ci, shp in enumerate(convs):
                out.append(("%s.conv%d" % (p, ci + 1), shp, "conv"))
            for ci, shp in enumerate(convs):
                out.append(("%s.bn%d" % (p, ci + 1), shp[0], "bn"))
            if b == 0 and (stride != 1 or inplanes != planes * exp):
                out.append((p + ".downsample.0", (planes * exp, inplanes, 1, 1), "conv"))
                out.append((p + ".downsample.1", planes * exp, "bn"))
        inplanes = planes * exp
    return out


def make_resnet_state(depth, num_stages=4, strides=(1, 2, 2, 2), generator=None):
    """Random-init weights as ``ResNet.init_weights(pretrained=None)`` defines them
    (models/backbone/resnet.py:244-249; kaiming_init models/utils/inits.py:33-46 with
    mode='fan_out', nonlinearity='relu', normal).  Note: values are *statistically* equal to the
    reference's, not RNG-stream-identical (the reference first runs nn.Conv2d's default init); tests
    that need identical weights copy the reference ``state_dict`` instead."""
    sd = OrderedDict()
    for name, shp, kind in resnet_param_shapes(depth, num_stages, strides):
        if kind == "conv":
            fan_out = shp[0] * shp[2] * shp[3]
            std = (2.0 / fan_out) ** 0.5
            sd[name + ".weight"] = torch.randn(shp, generator=generator) * std
        else:
            _bn_entries(sd, name, shp)
    return sd


def fpn_level_range(num_ins, start_level=0, end_level=-1):
    return start_level, (num_ins if end_level == -1 else end_level)


def make_fpn_state(in_channels, out_channels, num_outs, start_level=0, end_level=-1,
                   add_extra_convs=False, generator=None):
    """Xavier-uniform weights / zero bias (models/necks/fpn.py:80-86, inits.py:11-18).  Key order:
    lateral_convs.* first, then fpn_convs.* (ModuleList registration order, fpn.py:40-78)."""
    lo, hi = fpn_level_range(len(in_channels), start_level, end_level)
    lat, out = OrderedDict(), OrderedDict()

    def xavier(shape):
        fan_in = shape[1] * shape[2] * shape[3]
        fan_out = shape[0] * shape[2] * shape[3]
        bound = (6.0 / (fan_in + fan_out)) ** 0.5
        return (torch.rand(shape, generator=generator) * 2 - 1) * bound

    for j, i in enumerate(range(lo, hi)):
        lat["lateral_convs.%d.conv.weight" % j] = xavier((out_channels, in_channels[i], 1, 1))
        lat["lateral_convs.%d.conv.bias" % j] = torch.zeros(out_channels)
        out["fpn_convs.%d.conv.weight" % j] = xavier((out_channels, out_channels, 3, 3))
        out["fpn_convs.%d.conv.bias" % j] = torch.zeros(out_channels)
    extra = num_outs - hi + lo
    if add_extra_convs and extra >= 1:
        for e in range(extra):
            cin = in_channels[hi - 1] if e == 0 else out_channels
            j = hi - lo + e
            out["fpn_convs.%d.conv.weight" % j] = xavier((out_channels, cin, 3, 3))
            out["fpn_convs.%d.conv.bias" % j] = torch.zeros(out_channels)
    sd = OrderedDict()
    sd.update(lat)
    sd.update(out)
    return sd


def randomize_bn_stats(sd, generator=None):
    """Give every BN non-trivial gamma/beta/mean/var (the reference init makes BN ~identity, which
    would hide BN-fold bugs)."""
    for k in list(sd.keys()):
        if k.endswith(".running_var"):
            p = k[: -len(".running_var")]
            ch = sd[k].numel()
            sd[p + ".weight"] = 0.5 + torch.rand(ch, generator=generator)
            sd[p + ".bias"] = 0.2 * torch.randn(ch, generator=generator)
            sd[p + ".running_mean"] = 0.2 * torch.randn(ch, generator=generator)
            sd[p + ".running_var"] = 0.5 + torch.rand(ch, generator=generator)
    return sd


def randomize_gn_affine(sd, generator=None):
    """Non-trivial gamma / beta for every GroupNorm of a use_gn=True state (init makes them 1 / 0): keys ``*.weight``
    of rank 1 whose module has no running statistics."""
    for k in list(sd.keys()):
        if k.endswith(".weight") and sd[k].dim() == 1 and (k[:-len(".weight")] + ".running_var") not in sd:
            p = k[:-len(".weight")]
            ch = sd[k].numel()
            sd[p + ".weight"] = 0.5 + torch.rand(ch, generator=generator)
            sd[p + ".bias"] = 0.2 * torch.randn(ch, generator=generator)
    return sd


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------

GN_GROUPS = 32  # get_group_gn (models/utils/layers.py:138-154): num_groups = 32 whatever the width
GN_EPS = 1e-5   # nn.GroupNorm default


def _bn(sd, p, x):
    if (p + ".running_mean") not in sd:
        # use_gn=True: norm_layer -> nn.GroupNorm(get_group_gn(planes), planes) (layers.py:50-54); no running statistics
        return F.group_norm(x, GN_GROUPS, sd[p + ".weight"], sd[p + ".bias"], GN_EPS)
    # eval-mode BatchNorm2d == F.batch_norm(training=False): aten::native_batch_norm, as nn.BatchNorm2d
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], False, 0.1, BN_EPS)


def _block(sd, p, x, kind, stride, dilation):
    has_ds = (p + ".downsample.0.weight") in sd
    nm = ".gn" if (p + ".gn1.weight") in sd else ".bn"  # norm_names with use_gn=True (resnet.py:31,85-86)
    if kind == "bottleneck":  # resnet.py:97-119 ; stride on the 3x3 (:75)
        out = F.conv2d(x, sd[p + ".conv1.weight"])
        out = F.relu_(_bn(sd, p + nm + "1", out))
        # groups > 1 for ResNeXt (models/backbone/resnext.py:84-87): inferred from the parameter's shape
        w2 = sd[p + ".conv2.weight"]
        out = F.conv2d(out, w2, None, stride, dilation, dilation, out.shape[1] // w2.shape[1])
        out = F.relu_(_bn(sd, p + nm + "2", out))
        out = F.conv2d(out, sd[p + ".conv3.weight"])
        out = _bn(sd, p + nm + "3", out)
    else:  # resnet.py:42-59 ; second 3x3 has dilation 1 / pad 1 (:23-24)
        out = F.conv2d(x, sd[p + ".conv1.weight"], None, stride, dilation, dilation)
        out = F.relu_(_bn(sd, p + nm + "1", out))
        out = F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1, 1)
        out = _bn(sd, p + nm + "2", out)
    residual = x
    if has_ds:  # resnet.py:129-136
        residual = F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride)
        residual = _bn(sd, p + ".downsample.1", residual)
    out += residual
    return F.relu_(out)


def resnet_forward(sd, x, depth, num_stages=4, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1),
                   out_indices=(0, 1, 2, 3)):
    """ResNet.forward in eval/frozen-BN mode (models/backbone/resnet.py:253-268)."""
    kind, counts = ARCH[depth]
    x = F.conv2d(x, sd["conv1.weight"], None, 2, 3)
    x = F.relu_(_bn(sd, "gn1" if "gn1.weight" in sd else "bn1", x))  # norm_name (resnet.py:215)
    x = F.max_pool2d(x, 3, 2, 1)
    outs = []
    for li, nblocks in enumerate(counts[:num_stages]):
        for b in range(nblocks):
            x = _block(sd, "layer%d.%d" % (li + 1, b), x, kind,
                       strides[li] if b == 0 else 1, dilations[li])
        if li in out_indices:
            outs.append(x)
    return outs[0] if len(outs) == 1 else tuple(outs)


def _cm(sd, prefix, x, stride=1, pad=0):
    """ConvModule.forward with activate_last=True and no activation (layers.py:120-127): conv (+bias), then the
    norm layer if the module has one (eval-mode BatchNorm2d, or GroupNorm with use_gn=True)."""
    y = F.conv2d(x, sd[prefix + ".conv.weight"], sd.get(prefix + ".conv.bias"), stride, pad)
    if (prefix + ".norm.weight") in sd:
        y = _bn(sd, prefix + ".norm", y)
    return y


def fpn_forward(sd, inputs, in_channels, out_channels, num_outs, start_level=0, end_level=-1,
                add_extra_convs=False):
    """FPN.forward (models/necks/fpn.py:88-125), normalize=None (conv + bias only)."""
    assert len(inputs) == len(in_channels)
    lo, hi = fpn_level_range(len(in_channels), start_level, end_level)
    n = hi - lo
    lats = [_cm(sd, "lateral_convs.%d" % j, inputs[lo + j]) for j in range(n)]
    for j in range(n - 1, 0, -1):
        lats[j - 1] += F.interpolate(lats[j], scale_factor=2, mode="nearest")
    outs = [_cm(sd, "fpn_convs.%d" % j, lats[j], 1, 1) for j in range(n)]
    if num_outs > len(outs):
        if not add_extra_convs:
            for _ in range(num_outs - n):
                outs.append(F.max_pool2d(outs[-1], 1, stride=2))
        else:
            outs.append(_cm(sd, "fpn_convs.%d" % n, inputs[hi - 1], 2, 1))
            for j in range(n + 1, num_outs):
                outs.append(_cm(sd, "fpn_convs.%d" % j, F.relu(outs[-1], inplace=True), 2, 1))
    return tuple(outs)


def pafpn_forward(sd, inputs, in_channels, out_channels, num_outs, start_level=0, end_level=-1,
                  add_extra_convs=False, activation=None):
    """PAFPN.forward (models/necks/pafpn.py:108-148), normalize=None: the FPN pyramid P, then the
    bottom-up path N_i = pa_convs2[i-1](P_i + pa_convs1[i-1](N_{i-1})) (:131-134); `activation` is the
    ConvModule activation of the pa convs (None, 'relu' or 'relu6', applied to each conv's output)."""
    assert len(inputs) == len(in_channels)
    lo, hi = fpn_level_range(len(in_channels), start_level, end_level)
    n = hi - lo
    act = {None: (lambda t: t), "relu": (lambda t: F.relu(t, inplace=True)),
           "relu6": (lambda t: F.relu6(t, inplace=True))}[activation]  # layers.py:114-119
    lats = [_cm(sd, "lateral_convs.%d" % j, inputs[lo + j]) for j in range(n)]
    for j in range(n - 1, 0, -1):
        lats[j - 1] += F.interpolate(lats[j], scale_factor=2, mode="nearest")
    outs = [_cm(sd, "fpn_convs.%d" % j, lats[j], 1, 1) for j in range(n)]
    for j in range(1, n):
        down = act(_cm(sd, "pa_convs1.%d" % (j - 1), outs[j - 1], 2, 1))
        outs[j] = act(_cm(sd, "pa_convs2.%d" % (j - 1), outs[j] + down, 1, 1))
    if num_outs > len(outs):
        if not add_extra_convs:
            for _ in range(num_outs - n):
                outs.append(F.max_pool2d(outs[-1], 1, stride=2))
        else:
            outs.append(_cm(sd, "fpn_convs.%d" % n, inputs[hi - 1], 2, 1))
            for j in range(n + 1, num_outs):
                outs.append(_cm(sd, "fpn_convs.%d" % j, F.relu(outs[-1], inplace=True), 2, 1))
    return tuple(outs)


def resnet_fpn_forward(bb_sd, neck_sd, x, depth, out_channels=256, num_outs=5):
    kind, _ = ARCH[depth]
    in_ch = [64 * 2 ** i * EXPANSION[kind] for i in range(4)]
    feats = resnet_forward(bb_sd, x, depth)
    return feats, fpn_forward(neck_sd, feats, in_ch, out_channels, num_outs)


# ----------------------------------------------------------------------------------------------
# precision emulation of the CUDA path (SURVEY.md Appendix D "emulation A", Appendix E-5): bf16
# conv operands, fp32 accumulate/epilogue, one bf16 rounding per stored tensor.  Used by tests to
# separate "kernel bug" from "bf16 is bf16".
# ----------------------------------------------------------------------------------------------

def _r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def resnet_fpn_forward_bf16_emulated(bb_sd, neck_sd, x, depth, out_channels=256, num_outs=5,
                                     weight_dtype=torch.float16):
    """weight_dtype: 16-bit format the kernels keep conv weights in (fp16 by default; pass
    torch.bfloat16 for SURVEY.md's original 'emulation A')."""
    kind, counts = ARCH[depth]
    bb_sd = {k: (v.to(weight_dtype).float() if k.endswith("weight") and v.dim() == 4 else v)
             for k, v in bb_sd.items()}
    neck_sd = {k: (v.to(weight_dtype).float() if k.endswith("weight") and v.dim() == 4 else v)
               for k, v in neck_sd.items()}

    def cbn(inp, wkey, bnp, stride=1, pad=0, relu=False, res=None):
        y = F.conv2d(inp, bb_sd[wkey], None, stride, pad)
        scale = bb_sd[bnp + ".weight"] / torch.sqrt(bb_sd[bnp + ".running_var"] + BN_EPS)
        shift = bb_sd[bnp + ".bias"] - bb_sd[bnp + ".running_mean"] * scale
        y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        if res is not None:
            y = y + res
        if relu:
            y = F.relu(y)
        return _r(y)

    x = cbn(_r(x), "conv1.weight", "bn1", 2, 3, relu=True)
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li, nblocks in enumerate(counts):
        for b in range(nblocks):
            p = "layer%d.%d" % (li + 1, b)
            s = 2 if (b == 0 and li > 0) else 1
            res = x
            if (p + ".downsample.0.weight") in bb_sd:
                res = cbn(x, p + ".downsample.0.weight", p + ".downsample.1", s, 0)
            if kind == "bottleneck":
                o = cbn(x, p + ".conv1.weight", p + ".bn1", 1, 0, relu=True)
                o = cbn(o, p + ".conv2.weight", p + ".bn2", s, 1, relu=True)
                x = cbn(o, p + ".conv3.weight", p + ".bn3", 1, 0, relu=True, res=res)
            else:
                o = cbn(x, p + ".conv1.weight", p + ".bn1", s, 1, relu=True)
                x = cbn(o, p + ".conv2.weight", p + ".bn2", 1, 1, relu=True, res=res)
        feats.append(x)
    n = len(feats)
    lats = [None] * n
    for j in range(n - 1, -1, -1):
        y = F.conv2d(feats[j], neck_sd["lateral_convs.%d.conv.weight" % j],
                     neck_sd["lateral_convs.%d.conv.bias" % j])
        if j < n - 1:
            y = y + F.interpolate(lats[j + 1], scale_factor=2, mode="nearest")
        lats[j] = _r(y)
    outs = [_r(F.conv2d(lats[j], neck_sd["fpn_convs.%d.conv.weight" % j],
                        neck_sd["fpn_convs.%d.conv.bias" % j], 1, 1)) for j in range(n)]
    for _ in range(num_outs - n):
        outs.append(F.max_pool2d(outs[-1], 1, stride=2))
    return tuple(feats), tuple(outs)


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 in fp64 (b = oracle)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def conv_flops(depth, h, w, with_fpn=True, out_channels=256):
    """2*M*N*K over convs, un-padded dims, per image."""
    def o(v, k, s, p):
        return (v + 2 * p - (k - 1) - 1) // s + 1
    kind, counts = ARCH[depth]
    exp = EXPANSION[kind]
    fl = 0
    h, w = o(h, 7, 2, 3), o(w, 7, 2, 3)
    fl += 2 * h * w * 64 * 147
    h, w = o(h, 3, 2, 1), o(w, 3, 2, 1)
    inpl = 64
    feats = []
    for li, nb in enumerate(counts):
        pl = 64 * 2 ** li
        for b in range(nb):
            s = 2 if (b == 0 and li > 0) else 1
            cin = inpl if b == 0 else pl * exp
            ho, wo = o(h, 3, s, 1), o(w, 3, s, 1)
            if kind == "bottleneck":
                fl += 2 * h * w * pl * cin + 2 * ho * wo * pl * pl * 9 + 2 * ho * wo * pl * exp * pl
            else:
                fl += 2 * ho * wo * pl * cin * 9 + 2 * ho * wo * pl * pl * 9
            if b == 0 and (s != 1 or inpl != pl * exp):
                fl += 2 * ho * wo * pl * exp * inpl
            h, w = ho, wo
        inpl = pl * exp
        feats.append((inpl, h, w))
    bb = fl
    if with_fpn:
        for c, fh, fw in feats:
            fl += 2 * fh * fw * out_channels * c + 2 * fh * fw * out_channels * out_channels * 9
    return fl, bb
