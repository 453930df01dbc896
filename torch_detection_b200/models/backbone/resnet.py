"""ResNet backbone, B200 execution -- drop-in for reference ``models/backbone/resnet.py``.

Same constructor arguments, attributes, ``state_dict`` keys/shapes, ``init_weights`` and
``train`` semantics as the reference ``ResNet`` (resnet.py:158-294); ``forward`` takes the same
``(N,3,H,W)`` batch and returns the same stage-feature tuple (a bare tensor for a single
``out_indices`` entry, resnet.py:265-268).  What differs is *how* forward runs: the module compiles
its topology once per input shape into a list of ``tdet_op`` descriptors (stem, max-pool and one
fused conv+BN(+residual)(+ReLU) op per convolution) and hands it to ``libtdet_b200.so``; the
``nn.Conv2d`` / ``nn.BatchNorm2d`` children only hold the fp32 master parameters.

Outputs are dense NHWC bf16 buffers exposed as logical-NCHW ``channels_last`` tensors (fp32 inputs
get fp32 views-by-copy of the same values).  Deliberate deviations from the reference, all listed in
SURVEY.md Appendix G: ``train()`` returns ``self`` (reference returns None, F4) and implements the
*intended* stage freezing (the reference raises AttributeError at resnet.py:288, F3).
Unsupported on this path -> ``NotImplementedError``: ``use_gn=True``, batch-statistics BatchNorm
(a BN child in training mode), CPU tensors.  There is no fallback.
"""
import logging
import os

import torch
import torch.nn as nn

from ... import engine
from ...registry import BACKBONES
from ..utils import (conv1x1_group, conv3x3_group, conv7x7_group, norm_layer, kaiming_init,
                     constant_init, load_checkpoint)


class _ResidualUnit(nn.Module):
    """Parameter container for one residual unit.  Child names follow the reference so that
    ``state_dict`` keys match (``convK``, ``bnK``, ``downsample.{0,1}``)."""
    expansion = 1
    kernel_sizes = ()

    def __init__(self, inplanes, planes, stride=1, dilation=1, use_gn=False, downsample=None):
        super(_ResidualUnit, self).__init__()
        widths = self._widths(inplanes, planes)
        convs = []
        for idx, (k, (cin, cout)) in enumerate(zip(self.kernel_sizes, widths)):
            strided = idx == self.strided_conv
            if k == 1:
                convs.append(conv1x1_group(cin, cout))
            else:
                convs.append(conv3x3_group(cin, cout, stride if strided else 1,
                                           dilation if (strided or self.all_dilated) else 1))
        for idx, conv in enumerate(convs):
            self.add_module("conv%d" % (idx + 1), conv)
        self.norm_names = ["bn%d" % (i + 1) for i in range(len(convs))]
        for name, (_, cout) in zip(self.norm_names, widths):
            self.add_module(name, norm_layer(cout, use_gn))
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self.dilation = dilation
        self.use_gn = use_gn

    def forward(self, x):
        raise NotImplementedError("residual units are executed by the owning ResNet's plan")


class BasicBlock(_ResidualUnit):
    """3x3(stride, dil) -> BN -> ReLU -> 3x3 -> BN -> (+residual) -> ReLU   (resnet.py:9-59)"""
    expansion = 1
    kernel_sizes = (3, 3)
    strided_conv = 0
    all_dilated = False

    @staticmethod
    def _widths(inplanes, planes):
        return [(inplanes, planes), (planes, planes)]


class Bottleneck(_ResidualUnit):
    """1x1 -> BN -> ReLU -> 3x3(stride, dil) -> BN -> ReLU -> 1x1 -> BN -> (+residual) -> ReLU
    (resnet.py:62-119; the stride sits on the 3x3, :75)"""
    expansion = 4
    kernel_sizes = (1, 3, 1)
    strided_conv = 1
    all_dilated = False

    @staticmethod
    def _widths(inplanes, planes):
        return [(inplanes, planes), (planes, planes), (planes, planes * 4)]


def _make_res_layer(block, inplanes, planes, blocks, stride=1, dilation=1, use_gn=False):
    """One stage (resnet.py:122-155).  The projection shortcut is constructed before the first unit,
    as in the reference, so parameter creation consumes the RNG in the same order."""
    shortcut = None
    if stride != 1 or inplanes != planes * block.expansion:
        shortcut = nn.Sequential(conv1x1_group(inplanes, planes * block.expansion, stride=stride),
                                 norm_layer(planes * block.expansion, use_gn=use_gn))
    units = [block(inplanes, planes, stride=stride, dilation=dilation, use_gn=use_gn,
                   downsample=shortcut)]
    for _ in range(1, blocks):
        units.append(block(planes * block.expansion, planes, stride=1, dilation=dilation,
                           use_gn=use_gn))
    return nn.Sequential(*units)


@BACKBONES.register_module
class ResNet(nn.Module):
    """ResNet-{18,34,50,101,152} feature extractor.  See the module docstring."""

    arch_settings = {
        18: (BasicBlock, (2, 2, 2, 2)),
        34: (BasicBlock, (3, 4, 6, 3)),
        50: (Bottleneck, (3, 4, 6, 3)),
        101: (Bottleneck, (3, 4, 23, 3)),
        152: (Bottleneck, (3, 8, 36, 3)),
    }

    def __init__(self, depth, num_stages=4, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1),
                 out_indices=(0, 1, 2, 3), frozen_stages=-1, use_gn=False, bn_eval=True,
                 bn_frozen=False):
        super(ResNet, self).__init__()
        if depth not in self.arch_settings:
            raise KeyError("invalid depth {} for resnet".format(depth))
        assert 1 <= num_stages <= 4
        block, stage_blocks = self.arch_settings[depth]
        stage_blocks = stage_blocks[:num_stages]
        assert len(strides) == len(dilations) == num_stages
        assert max(out_indices) < num_stages

        self.depth = depth
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
        self.norm_name = "bn1"
        self.add_module(self.norm_name, norm_layer(64, use_gn))
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)

        self.res_layers = []
        for i, num_blocks in enumerate(stage_blocks):
            planes = 64 * 2 ** i
            stage = _make_res_layer(block, self.inplanes, planes, num_blocks, stride=strides[i],
                                    dilation=dilations[i], use_gn=use_gn)
            self.inplanes = planes * block.expansion
            name = "layer{}".format(i + 1)
            self.add_module(name, stage)
            self.res_layers.append(name)
        self.feat_dim = block.expansion * 64 * 2 ** (len(stage_blocks) - 1)

        self._plans = {}
        self._operands = None
        self._operand_key = None

    # ------------------------------------------------------------------ reference API
    def init_weights(self, pretrained=None):
        if isinstance(pretrained, str):
            load_checkpoint(self, pretrained, strict=False, logger=logging.getLogger())
        elif pretrained is None:
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    kaiming_init(m)
                elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                    constant_init(m, 1)
        else:
            raise TypeError("pretrained must be a str or None")

    def train(self, mode=True):
        super(ResNet, self).train(mode)
        if not self.use_gn and self.bn_eval:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
                    if self.bn_frozen:
                        for p in m.parameters():
                            p.requires_grad = False
        if mode and self.frozen_stages >= 0:
            stem_norm = getattr(self, self.norm_name)
            for p in list(self.conv1.parameters()) + list(stem_norm.parameters()):
                p.requires_grad = False
            stem_norm.eval()
            for i in range(1, self.frozen_stages + 1):
                stage = getattr(self, "layer{}".format(i))
                stage.eval()
                for p in stage.parameters():
                    p.requires_grad = False
        return self

    # ------------------------------------------------------------------ B200 execution
    def _check_supported(self, x):
        engine.require_cuda(x, "ResNet input")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected an (N,3,H,W) batch, got %s" % (tuple(x.shape),))
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d) and m.training:
                raise NotImplementedError(
                    "batch-statistics BatchNorm is not supported on the B200 path (frozen/eval BN "
                    "only, as in the reference's configs): call .eval() or .train() with "
                    "bn_eval=True first")

    def _param_key(self, device):
        return (device,) + tuple((p.data_ptr(), p._version) for p in self.parameters()) + \
            tuple((b.data_ptr(), b._version) for b in self.buffers())

    def _get_operands(self, device):
        """Lazy cache of derived operands (packed weights per format, folded BN vectors, bound
        constants); dropped together with all plans whenever a parameter or buffer changes."""
        key = self._param_key(device)
        if self._operands is None or key != self._operand_key:
            self._operands = _OperandCache()
            self._operand_key = key
            self._plans = {}
        return self._operands

    def _segments(self, n):
        """[(stage indices, images per chunk)]: the stem travels with the first segment.  Early stages
        are HBM-bound layer by layer; running them a few images at a time keeps each block's tensors
        in the 126 MB L2 between consecutive kernels (DESIGN.md "L2-resident scheduling")."""
        nst = len(self.res_layers)
        spec = [int(v) for v in os.environ.get("TDET_CHUNKS", DEFAULT_CHUNKS).split(",") if v.strip()]
        segs = []
        for i in range(nst):
            c = spec[i] if i < len(spec) and spec[i] > 0 else n
            c = min(c, n)
            if segs and segs[-1][1] == c:
                segs[-1][0].append(i)
            else:
                segs.append(([i], c))
        return segs

    def _build_plan(self, x, cache):
        """Compiles the topology for one input geometry into tdet_ops.

        Internal activations are fp16 significands with a per-tensor power-of-two exponent chosen on
        the device (TDET_FLAG_SCALED_OUT); stage outputs are plain bf16 (they are what the module
        returns and what crosses chunk boundaries).  tcgen05 needs both MMA operands in one format,
        so each conv's weights are packed in its input's format."""
        n, _, h, w = x.shape
        dev = x.device
        internal = INTERNAL_DTYPE
        scaled = internal == torch.float16
        ops = []
        pool = _BufferPool(dev)
        segs = self._segments(n)
        max_chunks = max((n + c - 1) // c for _, c in segs)
        n_meta = 8 + (2 + 4 * sum(len(getattr(self, l)) for l in self.res_layers)) * max_chunks
        meta = engine.MetaArena(n_meta, dev)

        def new_act(shape, dtype):
            return engine.Act(pool.get(shape), shape, dtype, meta.new())

        def conv(name, module, bn, src, dst, residual=None, relu=True):
            k = module.kernel_size[0]
            wgt = cache.get((name, "w", src.dtype), lambda: engine.pack_conv_weight(module.weight, src.dtype))
            sc, sh = cache.get((name, "bn"), lambda: engine.fold_bn(bn))
            is_scaled = scaled and dst.dtype == torch.float16
            consts = cache.get((name, "consts", src.dtype),
                               lambda: engine.bound_consts(wgt, sc, sh)) if is_scaled else None
            ops.append(engine.op_conv(src, wgt, dst, k, k, module.stride[0], module.padding[0],
                                      module.dilation[0], scale=sc, shift=sh, residual=residual,
                                      relu=relu, consts=consts, scaled_out=is_scaled))

        # geometry of every stage output, and the full-batch bf16 tensors that hold them
        ho, wo = engine.conv_out(h, 7, 2, 3), engine.conv_out(w, 7, 2, 3)
        hq, wq = engine.conv_out(ho, 3, 2, 1), engine.conv_out(wo, 3, 2, 1)
        geo = []
        hh, ww = hq, wq
        for lname in self.res_layers:
            stage = getattr(self, lname)
            for unit in stage:
                hh = engine.conv_out(hh, 3, unit.stride, unit.dilation, unit.dilation)
                ww = engine.conv_out(ww, 3, unit.stride, unit.dilation, unit.dilation)
            last = stage[len(stage) - 1]
            cout = getattr(last, "conv%d" % len(last.kernel_sizes)).out_channels
            geo.append((hh, ww, cout))
        outs = []
        boundary = []
        for li, (sh_, sw_, sc_) in enumerate(geo):
            if li in self.out_indices:
                t = engine.nhwc_empty(n, sh_, sw_, sc_, dev)  # placeholder, re-bound at run time
                outs.append(t)
            else:
                t = pool.get((n, sh_, sw_, sc_))
            boundary.append((t, meta.new()))

        def boundary_act(li, i0, cn):
            t, m = boundary[li]
            sh_, sw_, sc_ = geo[li]
            return engine.Act(t, (cn, sh_, sw_, sc_), torch.bfloat16, m, offset=i0 * sh_ * sw_ * sc_)

        for stages, chunk in segs:
            for i0 in range(0, n, chunk):
                cn = min(chunk, n - i0)
                if stages[0] == 0:
                    staged = pool.get((cn,) + engine.stem_staging_dims(ho, wo) + (4,))
                    staged_meta = meta.new()
                    ops.append(engine.op_prep(x[i0:i0 + cn], staged, ho, wo, y_meta=staged_meta))
                    stem_out = new_act((cn, ho, wo, 64), internal)
                    stem_w = cache.get(("conv1", "w"), lambda: engine.pack_stem_weight(self.conv1.weight))
                    sc, sh = cache.get(("conv1", "bn"), lambda: engine.fold_bn(getattr(self, self.norm_name)))
                    stem_consts = cache.get(("conv1", "consts"),
                                            lambda: engine.bound_consts(stem_w, sc, sh)) if scaled else None
                    ops.append(engine.op_stem(cn, h, w, staged, stem_w, stem_out, sc, sh,
                                              x_meta=staged_meta, consts=stem_consts, scaled_out=scaled))
                    pool.release(staged)
                    # max-pool commutes with the (positive) per-tensor scale: metadata passes through
                    cur = engine.Act(pool.get((cn, hq, wq, 64)), (cn, hq, wq, 64), internal, stem_out.meta)
                    ops.append(engine.op_maxpool(stem_out, cur))
                    pool.release(stem_out.buf)
                    cur_pooled = True
                else:
                    cur = boundary_act(stages[0] - 1, i0, cn)
                    cur_pooled = False
                for li in stages:
                    lname = self.res_layers[li]
                    stage = getattr(self, lname)
                    for bi, unit in enumerate(stage):
                        pre = "%s.%d." % (lname, bi)
                        last = bi == len(stage) - 1
                        nb, hb, wb, _ = cur.shape
                        hn = engine.conv_out(hb, 3, unit.stride, unit.dilation, unit.dilation)
                        wn = engine.conv_out(wb, 3, unit.stride, unit.dilation, unit.dilation)
                        residual = cur
                        shortcut = None
                        if unit.downsample is not None:
                            cd = unit.downsample[0].out_channels
                            shortcut = new_act((nb, hn, wn, cd), internal)
                            conv(pre + "downsample", unit.downsample[0], unit.downsample[1], cur, shortcut,
                                 relu=False)
                            residual = shortcut
                        nconv = len(unit.kernel_sizes)
                        src = cur
                        temps = []
                        for ci, k in enumerate(unit.kernel_sizes):
                            module = getattr(unit, "conv%d" % (ci + 1))
                            oh = engine.conv_out(src.shape[1], k, module.stride[0], module.padding[0],
                                                 module.dilation[0])
                            ow = engine.conv_out(src.shape[2], k, module.stride[0], module.padding[0],
                                                 module.dilation[0])
                            final = ci == nconv - 1
                            if final and last:
                                dst = boundary_act(li, i0, cn)  # stage output: plain bf16
                            else:
                                dst = new_act((nb, oh, ow, module.out_channels), internal)
                                if not final:
                                    temps.append(dst)
                            conv(pre + "conv%d" % (ci + 1), module, getattr(unit, unit.norm_names[ci]), src,
                                 dst, residual=residual if final else None)
                            src = dst
                        for t in temps:
                            pool.release(t.buf)
                        if shortcut is not None:
                            pool.release(shortcut.buf)
                        if cur_pooled:
                            pool.release(cur.buf)
                        cur = src
                        cur_pooled = not last  # stage outputs live in boundary tensors, never pooled
        plan = engine.Plan(ops, [x] + outs, [cache, pool.all_buffers], dev, meta=meta)
        return plan, [tuple(o.shape) for o in outs]

    def forward(self, x):
        self._check_supported(x)
        cache = self._get_operands(x.device)
        key = (tuple(x.shape), x.dtype, tuple(x.stride()), x.device, INTERNAL_DTYPE)
        entry = self._plans.get(key)
        if entry is None:
            entry = self._build_plan(x, cache)
            self._plans[key] = entry
        plan, out_shapes = entry
        outs = [torch.empty(s, dtype=torch.bfloat16, device=x.device,
                            memory_format=torch.channels_last) for s in out_shapes]
        plan.run([x] + outs)
        self._last_run = (plan, [x] + outs)  # for profiling tools (bench.py)
        if x.dtype == torch.float32:
            outs = [_upcast(o) for o in outs]
        return outs[0] if len(outs) == 1 else tuple(outs)


# Storage format of the backbone's internal activations: float16 = fp16 significand + per-tensor
# power-of-two exponent (default; ~8x lower rounding error than bf16 at the same tensor-core rate),
# bfloat16 = plain bf16 everywhere (TDET_INTERNAL_DTYPE=bf16).
INTERNAL_DTYPE = torch.bfloat16 if os.environ.get("TDET_INTERNAL_DTYPE", "fp16").lower() in (
    "bf16", "bfloat16") else torch.float16

# Images per chunk for stage 1, 2, ... ("0" or missing = whole batch); TDET_CHUNKS overrides.
DEFAULT_CHUNKS = "0"


class _OperandCache(object):
    def __init__(self):
        self.store = {}

    def get(self, key, make):
        if key not in self.store:
            self.store[key] = make()
        return self.store[key]


def _upcast(t):
    """fp32 copy of a bf16 feature map that remembers its bf16 original, so a following B200 module
    (the FPN) consumes the bf16 buffer without a round trip."""
    f = t.float()
    f._tdet_bf16 = t
    return f


class _BufferPool(object):
    """Static activation arena for one plan: dense NHWC bf16 buffers, reused by byte size once their
    last consumer has been emitted (ops run in stream order, so reuse is safe)."""

    def __init__(self, device):
        self.device = device
        self.free = {}
        self.all_buffers = []

    def get(self, shape):
        """16-bit buffer with at least prod(shape) elements (dtype-agnostic: raw storage)."""
        numel = 1
        for s in shape:
            numel *= s
        bucket = self.free.get(numel)
        if bucket:
            return bucket.pop()
        t = torch.empty(numel, dtype=torch.bfloat16, device=self.device)
        self.all_buffers.append(t)
        return t

    def release(self, t):
        self.free.setdefault(t.numel(), []).append(t)
