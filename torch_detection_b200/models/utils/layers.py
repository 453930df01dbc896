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
    the raw conv; ``activate_last=False`` is refused (no neck of the reference uses it)."""

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
        raise NotImplementedError(
            "ConvModule is a parameter container here; the owning FPN executes it through the plan")
