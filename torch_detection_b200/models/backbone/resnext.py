"""ResNeXt backbone, B200 execution -- drop-in for reference ``models/backbone/resnext.py`` (SURVEY.md 8(f)
row f4: the grouped 3x3 is the first GEMM shape the ResNet path does not have).

Same constructor arguments (``depth, base_width, cardinality, ...``), attributes (``resX_layers``),
``state_dict`` keys/shapes and ``train`` semantics as the reference ``ResNeXt`` (resnext.py:178-330); the
execution machinery is the product ``ResNet``'s (plans of fused conv ops).  The grouped 3x3 of every
bottleneck (resnext.py:84-87, ``groups=cardinality`` with ``D = planes * base_width / 64`` channels per
group) runs on the dense implicit-GEMM kernel as a 64-channel *band*: the weights are packed
block-diagonally into the dense ``[Cout][3][3][Cin]`` layout, and each 64-wide tile of output channels
contracts only the 64 input channels of the same range (``tdet_op.groups``), i.e. 64/D times the grouped
FLOPs instead of the Cin/D times a naive dense conv would cost.  Requires D to divide 64 (true for the
usual 32x4d / 32x8d / 64x4d variants at every stage up to D = 64).

Only the bottleneck depths (50/101/152) are offered: the reference's basic-block ResNeXt is broken
(``_make_resX_layer`` passes ``base_width`` where ``ResNeXtBasicBlock`` expects ``cardinality``,
resnext.py:150-158 vs :14-21).  Training runs through the ResNet plan compiler: the grouped conv's data gradient
is a grouped conv over the block-diagonal dgrad operand, its weight gradient is accumulated dense and unpacked to the
block diagonal (``training.BackwardBuilder``).
"""
import math

import torch.nn as nn

from ... import engine
from ...registry import BACKBONES
from ..utils import conv7x7_group, norm_layer
from .resnet import ResNet, Bottleneck, _make_res_layer


def _resnext_bottleneck(base_width, cardinality):
    """Bottleneck subclass for one (base_width, cardinality): 1x1 -> grouped 3x3 (stride) -> 1x1."""

    class ResNeXtBottleneck(Bottleneck):
        groups = cardinality
        grouped_conv = 1

        @staticmethod
        def _widths(inplanes, planes):
            d = int(math.floor(planes * (base_width / 64.0)))
            mid = d * cardinality
            return [(inplanes, mid), (mid, mid), (mid, planes * 4)]

    ResNeXtBottleneck.base_width = base_width
    ResNeXtBottleneck.cardinality = cardinality
    return ResNeXtBottleneck


@BACKBONES.register_module
class ResNeXt(ResNet):
    """ResNeXt-{50,101,152} (``base_width`` x ``cardinality``) feature extractor."""

    arch_settings = {
        50: (3, 4, 6, 3),
        101: (3, 4, 23, 3),
        152: (3, 8, 36, 3),
    }

    def __init__(self, depth, base_width, cardinality, num_stages=4, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1),
                 out_indices=(0, 1, 2, 3), frozen_stages=-1, use_gn=False, bn_eval=True, bn_frozen=False):
        nn.Module.__init__(self)
        if depth in (18, 34):
            raise NotImplementedError("basic-block ResNeXt-%d is broken in the reference (resnext.py:150-158)" % depth)
        if depth not in self.arch_settings:
            raise KeyError("invalid depth {} for resnet".format(depth))
        assert 1 <= num_stages <= 4
        stage_blocks = self.arch_settings[depth][:num_stages]
        assert len(strides) == len(dilations) == num_stages
        assert max(out_indices) < num_stages
        block = _resnext_bottleneck(base_width, cardinality)

        self.depth = depth
        self.base_width = base_width
        self.cardinality = cardinality
        self.out_indices = out_indices
        self.frozen_stages = frozen_stages
        if not use_gn:
            self.bn_eval = bn_eval
            self.bn_frozen = bn_frozen
        self.use_gn = use_gn
        self.strides = tuple(strides)
        self.dilations = tuple(dilations)

        self.inplanes = 64
        self.conv1 = conv7x7_group(3, 64, stride=2)
        self.norm_name = "gn1" if use_gn else "bn1"
        self.add_module(self.norm_name, norm_layer(64, use_gn))
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)

        self.res_layers = []
        for i, num_blocks in enumerate(stage_blocks):
            planes = 64 * 2 ** i
            d = int(math.floor(planes * (base_width / 64.0)))
            if d < 1 or 64 % d:
                raise NotImplementedError(
                    "group width %d (stage %d) does not divide 64: outside the B200 grouped-conv band kernel" % (d, i + 1))
            stage = _make_res_layer(block, self.inplanes, planes, num_blocks, stride=strides[i],
                                    dilation=dilations[i], use_gn=use_gn)
            self.inplanes = planes * block.expansion
            name = "layer{}".format(i + 1)
            self.add_module(name, stage)
            self.res_layers.append(name)
        self.resX_layers = self.res_layers  # the reference's attribute name (resnext.py:246)
        self.feat_dim = block.expansion * 64 * 2 ** (len(stage_blocks) - 1)

        self._plans = engine.PlanCache()
        self._operands = None
        self._operand_key = None
