"""Parameter containers with the reference's factory names (models/utils/layers.py:6-135).

On the B200 path these modules only *hold* the fp32 master parameters under the reference's
``state_dict`` names; the arithmetic runs in libtdet_b200.so from packed bf16 copies.  They are real
``nn.Conv2d`` / ``nn.BatchNorm2d`` instances so checkpoints, optimisers and ``load_state_dict`` work
unchanged, and so that construction consumes the torch RNG exactly like the reference does.
"""
import warnings

import torch.nn as nn


def _conv(cin, cout, k, stride, padding, dilation, groups, bias):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=padding, dilation=dilation,
                     groups=groups, bias=bias)


def conv1x1_group(in_planes, out_planes, stride=1, groups=1):
    """1x1 conv, no bias (layers.py:6-17)."""
    return _conv(in_planes, out_planes, 1, stride, 0, 1, groups, False)


def conv3x3_group(in_planes, out_planes, stride=1, dilation=1, groups=1):
    """3x3 conv, padding == dilation, no bias (layers.py:20-32)."""
    return _conv(in_planes, out_planes, 3, stride, dilation, dilation, groups, False)


def conv7x7_group(in_planes, out_planes, stride=1, groups=1):
    """7x7 conv, padding 3, no bias (layers.py:35-47)."""
    return _conv(in_planes, out_planes, 7, stride, 3, 1, groups, False)


def get_group_gn(planes):
    """Number of GroupNorm groups for a width (layers.py:138-154): always 32."""
    num_groups = 32
    assert planes % num_groups == 0
    return num_groups


def norm_layer(planes, use_gn=False):
    """BatchNorm2d, or GroupNorm(32, planes) with use_gn=True (layers.py:50-54).  An eval-mode BatchNorm folds
    into the conv's epilogue; GroupNorm needs the statistics of the whole conv output, so it runs as two extra
    kernels after a raw conv (TDET_OP_GN_STATS / TDET_OP_GN_APPLY)."""
    if use_gn:
        return nn.GroupNorm(get_group_gn(planes), planes)
    return nn.BatchNorm2d(planes)


class ConvModule(nn.Module):
    """conv (+bias) (+BatchNorm) (+ReLU) parameter container used by the necks (layers.py:57-135).  The
    owning neck's plan executes it: an eval-mode BatchNorm (``normalize`` not None, ``use_gn=False``) is
    folded into the conv's fp32 epilogue, a GroupNorm (``use_gn=True``) runs as statistics + apply kernels after
    the raw conv; ``activate_last=False`` is refused (no neck of the reference uses it).  ``forward`` is a stand-alone
    inference call (one launch) for code that uses the class outside a neck."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, normalize=None, use_gn=False, activation=None,
                 activate_last=True):
        super(ConvModule, self).__init__()
        assert activation in (None, "relu", "relu6"), "Only ReLU and ReLU6 are supported"
        if not activate_last:
            raise NotImplementedError("ConvModule(activate_last=False) is not on the B200 neck path")
        self.with_norm = normalize is not None
        self.with_activation = activation is not None
        self.with_bias = bias
        self.activation = activation
        self.activate_last = activate_last
        if self.with_norm and self.with_bias:
            warnings.warn("ConvModule has norm and bias at the same time")
        self.conv = _conv(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                          bias)
        for attr in ("in_channels", "out_channels", "kernel_size", "stride", "padding", "dilation",
                     "groups"):
            setattr(self, attr, getattr(self.conv, attr))
        if self.with_norm:
            self.norm = norm_layer(out_channels, use_gn=use_gn)   # created after the conv, as in the reference
        if self.with_activation:
            self.activate = nn.ReLU6(inplace=True) if activation == "relu6" else nn.ReLU(inplace=True)

    def forward(self, x):
        """Stand-alone call, as heads built on the reference use it (layers.py:121-135); the necks do NOT come through
        here, they execute their ConvModules inside their own plans.  Inference only: ONE launch of the implicit-GEMM
        kernel with the bias / eval-mode BatchNorm and ReLU / ReLU6 in its epilogue, bf16 NHWC arithmetic, operands
        re-packed on every call (nothing is cached on this convenience path).  NCHW in, NCHW (channels_last memory) out,
        bf16 for a bf16 input and fp32 for an fp32 one."""
        import torch
        from ... import engine
        if torch.is_grad_enabled() and (x.requires_grad or (self.training and
                                                          any(p.requires_grad for p in self.parameters()))):
            raise NotImplementedError("stand-alone ConvModule training is not on the B200 path: the necks train "
                                      "through their plans; call it under torch.no_grad() / in eval mode")
        gn = self.with_norm and isinstance(self.norm, nn.GroupNorm)
        if self.with_norm and not gn and self.norm.training:
            raise NotImplementedError("batch-statistics BatchNorm is not on the B200 path (call .eval())")
        engine.require_cuda(x, "ConvModule input")
        conv = self.conv
        if conv.groups != 1 or conv.kernel_size[0] != conv.kernel_size[1] or conv.stride[0] != conv.stride[1] or \
                conv.padding[0] != conv.padding[1] or conv.dilation[0] != conv.dilation[1]:
            raise NotImplementedError("stand-alone ConvModule: dense convs with square geometry only")
        if x.dtype not in (torch.bfloat16, torch.float32):
            raise NotImplementedError("ConvModule input dtype %s (supported: float32, bfloat16)" % x.dtype)
        xb = x if (x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last)) else \
            x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wp = engine.pack_conv_weight(conv.weight, torch.bfloat16)
        k, st, pd, dl = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0]
        n, _, h, w = xb.shape
        oh, ow, co = engine.conv_out(h, k, st, pd, dl), engine.conv_out(w, k, st, pd, dl), conv.out_channels
        if gn:
            # GroupNorm's statistics span the whole conv output: raw conv (fp16 significands with a device-chosen
            # exponent) -> statistics -> normalise + affine + activation, as the necks' GroupNorm plans do (fpn.py)
            dev = x.device
            meta = engine.MetaArena(3, dev)
            src = engine.Act(xb, (n, h, w, xb.shape[1]), torch.bfloat16, meta.new())
            engine.run_op(engine.op_amax(engine.Act(xb, src.shape, torch.bfloat16), src.meta), dev)
            bias = conv.bias.detach().float().contiguous() if conv.bias is not None else None
            raw = engine.Act(torch.empty(n * oh * ow * co, dtype=torch.bfloat16, device=dev), (n, oh, ow, co),
                             torch.float16, meta.new())
            engine.run_op(engine.op_conv(src, wp, raw, k, k, st, pd, dl, shift=bias,
                                         consts=engine.bound_consts(wp, None, bias), scaled_out=True), dev)
            groups = self.norm.num_groups
            stats = torch.empty(engine.gn_stats_numel(n, groups), dtype=torch.float32, device=dev)
            y = engine.nhwc_empty(n, oh, ow, co, dev)
            engine.run_op(engine.op_gn_stats(raw, stats, groups), dev)
            engine.run_op(engine.op_gn_apply(raw, stats, groups, self.norm.weight.detach().float().contiguous(),
                                             self.norm.bias.detach().float().contiguous(), self.norm.eps,
                                             engine.Act(y, (n, oh, ow, co), torch.bfloat16, meta.new()),
                                             relu=self.activation == "relu", relu6=self.activation == "relu6"), dev)
            return y.float() if x.dtype == torch.float32 else y
        scale = shift = None
        if self.with_norm:
            scale, shift = engine.fold_bn(self.norm)
            if conv.bias is not None:
                shift = shift + conv.bias.detach().float() * scale   # BN(conv + b) = scale * conv + (shift + scale * b)
        elif conv.bias is not None:
            shift = conv.bias.detach().float().contiguous()
        y = engine.nhwc_empty(n, oh, ow, co, x.device)
        op = engine.op_conv(engine.act_of(xb), wp, engine.act_of(y), k, k, st, pd, dl, scale=scale, shift=shift,
                            relu=self.activation == "relu", relu6=self.activation == "relu6")
        engine.run_op(op, x.device)
        return y.float() if x.dtype == torch.float32 else y
