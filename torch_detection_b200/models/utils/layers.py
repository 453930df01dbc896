"""Parameter containers with the reference's factory names (models/utils/layers.py:6-135).

On the B200 path these modules only *hold* the fp32 master parameters under the reference's
``state_dict`` names; the arithmetic runs in libtdet_b200.so from packed bf16 copies.  They are real
``nn.Conv2d`` / ``nn.BatchNorm2d`` instances so checkpoints, optimisers and ``load_state_dict`` work
unchanged, and so that construction consumes the torch RNG exactly like the reference does.
"""
import warnings

import torch.nn as nn


def _conv(cin, cout, k, stride, padding, dilation, groups, bias):
    if groups != 1:
        raise NotImplementedError("grouped convolution is outside the B200 ResNet/FPN path")
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


def norm_layer(planes, use_gn=False):
    """BatchNorm2d (layers.py:50-54).  GroupNorm cannot be folded into a GEMM epilogue."""
    if use_gn:
        raise NotImplementedError("use_gn=True (GroupNorm) is not supported on the B200 path")
    return nn.BatchNorm2d(planes)


class ConvModule(nn.Module):
    """conv (+bias) container used by the neck (layers.py:57-135).  Only the configuration the FPN
    path uses is accepted: no norm, no activation (``normalize=None``)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, normalize=None, use_gn=False, activation=None,
                 activate_last=True):
        super(ConvModule, self).__init__()
        if normalize is not None or use_gn:
            raise NotImplementedError("ConvModule with a norm layer is not on the B200 FPN path")
        if activation is not None:
            raise NotImplementedError("ConvModule activation is not on the B200 FPN path")
        self.with_norm = False
        self.with_activation = False
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

    def forward(self, x):
        raise NotImplementedError(
            "ConvModule is a parameter container here; the owning FPN executes it through the plan")
