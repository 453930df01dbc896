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
``use_gn=True`` (GroupNorm) runs inference through ``_build_plan_gn`` (raw conv -> statistics -> apply).
Unsupported on this path -> ``NotImplementedError``: batch-statistics BatchNorm (a BN child in training
mode), training through GroupNorm, CPU tensors.  There is no fallback.
"""
import logging
import os

import torch
import torch.nn as nn

from ... import engine, training
from ...registry import BACKBONES
from ..utils import (conv1x1_group, conv3x3_group, conv7x7_group, norm_layer, kaiming_init,
                     constant_init, load_checkpoint)


class _ResidualUnit(nn.Module):
    """Parameter container for one residual unit.  Child names follow the reference so that
    ``state_dict`` keys match (``convK``, ``bnK``, ``downsample.{0,1}``)."""
    expansion = 1
    kernel_sizes = ()
    groups = 1          # ResNeXt: cardinality of the grouped 3x3 (resnext.py:84-87)
    grouped_conv = -1   # index of the grouped conv

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
                                           dilation if (strided or self.all_dilated) else 1,
                                           groups=self.groups if idx == self.grouped_conv else 1))
        for idx, conv in enumerate(convs):
            self.add_module("conv%d" % (idx + 1), conv)
        # resnet.py:31,85-86
        self.norm_names = [("gn%d" if use_gn else "bn%d") % (i + 1) for i in range(len(convs))]
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
        self.norm_name = "gn1" if use_gn else "bn1"  # resnet.py:215
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

        self._plans = engine.PlanCache()
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
    def set_input_transform(self, img_means=(0., 0., 0.), img_stds=(1., 1., 1.), size_divisor=None):
        """Fold the data layer's per-channel normalisation and pad-to-size-divisor (reference
        ``ImageTransforms`` steps 2 and 5, datasets/dataset_transforms.py:29-44) into
        the stem's loader kernel: ``forward`` then takes the RAW resized batch -- logical (N,3,h,w), uint8 /
        fp32 / bf16, any strides (an HWC uint8 batch viewed with ``.permute(0, 3, 1, 2)`` is zero-copy) --
        and the returned feature maps are those of the normalised image zero-padded to multiples of
        ``size_divisor``.  ``set_input_transform(None)`` removes it."""
        if img_means is None:
            self._input_tf = None
        else:
            means, stds = tuple(float(v) for v in img_means), tuple(float(v) for v in img_stds)
            assert len(means) == 3 and len(stds) == 3 and all(s != 0 for s in stds)
            self._input_tf = (means, stds, int(size_divisor) if size_divisor else None)
        self._plans = engine.PlanCache()

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

    def invalidate_operands(self):
        """Re-derive every packed weight / folded BatchNorm operand at the next forward.  Needed only after
        in-place parameter updates made through ``.data`` (they do not bump ``tensor._version``, which is how
        changes are detected otherwise: optimizer steps through ``torch.optim`` and ``load_state_dict`` are seen)."""
        self._operands_stale = True

    def _get_operands(self, device):
        """Lazy cache of derived operands (packed weights per format, folded BN vectors, bound
        constants).  When a parameter or buffer changes the affected entries are re-derived in place
        (engine.OperandCache.refresh): device pointers, hence the compiled plans, stay valid."""
        if self._operands is None or self._operand_key != device:
            self._operands = engine.OperandCache()
            self._operand_key = device
            self._plans = engine.PlanCache()
        else:
            force = getattr(self, "_operands_stale", False) or (engine.REFRESH_EVERY_STEP and self.training)
            self._operands.refresh(force=force)
        self._operands_stale = False
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

    def _build_plan(self, x, cache, train_from=None):
        """Compiles the topology for one input geometry into tdet_ops.

        train_from (training path): index of the first trainable stage.  Nothing from that stage on is
        recycled, stage outputs are kept in the internal format too (the returned bf16 tensors are
        converted copies: the wgrad MMA needs the saved input and the gradient in ONE format), and the
        per-block records the backward plan needs are returned.

        Internal activations are fp16 significands with a per-tensor power-of-two exponent chosen on
        the device (TDET_FLAG_SCALED_OUT); stage outputs are plain bf16 (they are what the module
        returns and what crosses chunk boundaries).  tcgen05 needs both MMA operands in one format,
        so each conv's weights are packed in its input's format."""
        n, _, h_in, w_in = x.shape
        dev = x.device
        tf = getattr(self, "_input_tf", None)
        h, w = h_in, w_in
        tf_scale = tf_shift = None
        if tf is not None:
            means, stds, div = tf
            if div:
                h, w = (h_in + div - 1) // div * div, (w_in + div - 1) // div * div
            tf_scale, tf_shift = cache.get(("input_tf", tf), lambda out: (
                torch.tensor([1.0 / s for s in stds], dtype=torch.float32, device=dev),
                torch.tensor([-m / s for m, s in zip(means, stds)], dtype=torch.float32, device=dev)))
        train = train_from is not None
        # fp32 input in eval mode = the fp32-I/O mode: split-precision tensors (bf16 hi|lo pairs, 2x the
        # channels in memory, every conv a 3-pass hi/lo GEMM) and fp32 outputs within 1e-4 of the fp32 reference
        split = (x.dtype == torch.float32) and not train and FP32_IO_SPLIT
        cs = 2 if split else 1
        internal = torch.bfloat16 if split else INTERNAL_DTYPE
        scaled = internal == torch.float16
        ops = []
        pool = _BufferPool(dev)
        segs = [(list(range(len(self.res_layers))), n)] if (train or split) else self._segments(n)
        records = []
        stem_rec = None
        max_chunks = max((n + c - 1) // c for _, c in segs)
        n_meta = 16 + (2 + 4 * sum(len(getattr(self, l)) for l in self.res_layers)) * max_chunks
        meta = engine.MetaArena(n_meta, dev)

        def new_act(shape, dtype):
            return engine.Act(pool.get(shape[:3] + (shape[3] * cs,)), shape, dtype, None if split else meta.new())

        def conv(name, module, bn, src, dst, residual=None, relu=True, shortcut=None):
            """shortcut = (name, conv, bn, input Act): the block's projection shortcut, contracted by this launch
            (TDET_FLAG_DUAL) instead of a launch of its own whose output comes back as `residual`."""
            k = module.kernel_size[0]
            if shortcut is not None:
                # one GEMM over [src | shortcut input] with the K-concatenated weights [scale*W | scale_s*W_s]: the
                # two BatchNorm scales are folded into the rows, the shifts are summed
                sname, smod, sbn, sx = shortcut
                sc, sh = cache.get((name, "bn"), lambda out: engine.fold_bn(bn, out=out), deps=_bn_deps(bn))
                sc2, sh2 = cache.get((sname, "bn"), lambda out: engine.fold_bn(sbn, out=out), deps=_bn_deps(sbn))
                deps = (module.weight, smod.weight) + _bn_deps(bn) + _bn_deps(sbn)
                wgt = cache.get((name, "wdual", src.dtype),
                                lambda out: engine.pack_dual_weight(module.weight, sc, smod.weight, sc2, src.dtype, out=out),
                                deps=deps)
                shs = cache.get((name, "shift_dual"), lambda out: _sum_into(sh, sh2, out), deps=_bn_deps(bn) + _bn_deps(sbn))
                is_scaled = scaled and dst.dtype == torch.float16
                consts = cache.get((name, "consts_dual", src.dtype),
                                   lambda out: engine.bound_consts(wgt, None, shs, out=out), deps=deps) if is_scaled else None
                ops.append(engine.op_conv(src, wgt, dst, 1, 1, 1, 0, 1, shift=shs, relu=relu, consts=consts,
                                          scaled_out=is_scaled, dual=(sx, smod.stride[0])))
                return
            if split:
                if module.groups > 1:
                    raise NotImplementedError("grouped convolutions have no split-precision (fp32-I/O) path yet")
                wgt = cache.get((name, "wsplit"), lambda out: engine.pack_conv_weight_split(module.weight, out=out),
                                deps=(module.weight,))
            elif module.groups > 1:
                wgt = cache.get((name, "w", src.dtype),
                                lambda out: engine.pack_grouped_conv_weight(module.weight, module.groups, src.dtype,
                                                                            out=out),
                                deps=(module.weight,))
            else:
                wgt = cache.get((name, "w", src.dtype),
                                lambda out: engine.pack_conv_weight(module.weight, src.dtype, out=out),
                                deps=(module.weight,))
            sc, sh = cache.get((name, "bn"), lambda out: engine.fold_bn(bn, out=out), deps=_bn_deps(bn))
            is_scaled = scaled and dst.dtype == torch.float16
            consts = cache.get((name, "consts", src.dtype),
                               lambda out: engine.bound_consts(wgt, sc, sh, out=out),
                               deps=(module.weight,) + _bn_deps(bn)) if is_scaled else None
            ops.append(engine.op_conv(src, wgt, dst, k, k, module.stride[0], module.padding[0],
                                      module.dilation[0], scale=sc, shift=sh, residual=residual,
                                      relu=relu, consts=consts, scaled_out=is_scaled, groups=module.groups,
                                      split=split))

        def tail(pre, unit, z1, xin, dst, head, head_pre):
            """Emits the fused tail of `unit` (input z1 = its conv1 output, residual xin, output dst) and returns the
            conv1 output of the following block `head` if that is computed by the same kernel (else None)."""
            def operands(name, module, bn, dtype):
                wgt = cache.get((name, "w", dtype), lambda out: engine.pack_conv_weight(module.weight, dtype, out=out),
                                deps=(module.weight,))
                sc, sh = cache.get((name, "bn"), lambda out: engine.fold_bn(bn, out=out), deps=_bn_deps(bn))
                consts = cache.get((name, "consts", dtype), lambda out: engine.bound_consts(wgt, sc, sh, out=out),
                                   deps=(module.weight,) + _bn_deps(bn)) if scaled else None
                return wgt, (sc, sh), consts
            w2, bn2, c2 = operands(pre + "conv2", unit.conv2, getattr(unit, unit.norm_names[1]), z1.dtype)
            w3, bn3, c3 = operands(pre + "conv3", unit.conv3, getattr(unit, unit.norm_names[2]), z1.dtype)
            nxt, y2 = None, None
            if head is not None:
                w1, bn1, c1 = operands(head_pre + "conv1", head.conv1, getattr(head, head.norm_names[0]), dst.dtype)
                y2 = new_act(z1.shape, internal)
                nxt = dict(w=w1, bn=bn1, y=y2, consts=c1, scaled_out=scaled and y2.dtype == torch.float16)
            ops.append(engine.op_bottleneck_tail(z1, w2, dst, xin, w3, bn2, bn3, consts2=c2, consts3=c3,
                                                 scaled_out=scaled and dst.dtype == torch.float16, nxt=nxt))
            return y2

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
        outs_f32 = []  # split mode: the returned fp32 tensors (hi + lo of the split stage outputs in `outs`)
        boundary = []
        for li, (sh_, sw_, sc_) in enumerate(geo):
            if li in self.out_indices:
                t = engine.nhwc_empty(n, sh_, sw_, sc_ * cs, dev)  # placeholder, re-bound at run time
                outs.append(t)
                if split:
                    outs_f32.append(engine.nhwc_empty(n, sh_, sw_, sc_, dev, torch.float32))
            else:
                t = pool.get((n, sh_, sw_, sc_ * cs))
            boundary.append((t, None if split else meta.new()))
        # training: stage outputs in the internal format (saved for backward, feed the next stage)
        cints = [engine.Act(pool.get((n, sh_, sw_, sc_)), (n, sh_, sw_, sc_), internal, meta.new())
                 for (sh_, sw_, sc_) in geo] if train else None

        def boundary_act(li, i0, cn):
            if train:
                return cints[li]
            t, m = boundary[li]
            sh_, sw_, sc_ = geo[li]
            return engine.Act(t, (cn, sh_, sw_, sc_), torch.bfloat16, m, offset=i0 * sh_ * sw_ * sc_)

        for stages, chunk in segs:
            for i0 in range(0, n, chunk):
                cn = min(chunk, n - i0)
                if stages[0] == 0:
                    staged = pool.get((cn * cs,) + engine.stem_staging_dims(ho, wo) + (4,))
                    staged_meta = None if split else meta.new()
                    ops.append(engine.op_prep(x[i0:i0 + cn], staged, ho, wo, y_meta=staged_meta, scale=tf_scale,
                                              shift=tf_shift, padded_hw=(h, w), split=split))
                    keep_stem = train and getattr(self, "_train_stem", False)
                    # the stem kernel max-pools in its epilogue unless the pre-pool activation is needed
                    # (stem training) or the tensors are bf16 pairs (fp32-I/O mode)
                    fuse_pool = FUSE_STEM_POOL and not split and not keep_stem
                    stem_out = None if fuse_pool else new_act((cn, ho, wo, 64), internal)
                    stem_bn = getattr(self, self.norm_name)
                    if split:
                        stem_w = cache.get(("conv1", "wsplit"),
                                           lambda out: engine.pack_stem_weight_split(self.conv1.weight, out=out),
                                           deps=(self.conv1.weight,))
                    else:
                        stem_w = cache.get(("conv1", "w"),
                                           lambda out: engine.pack_stem_weight(self.conv1.weight, out=out),
                                           deps=(self.conv1.weight,))
                    sc, sh = cache.get(("conv1", "bn"), lambda out: engine.fold_bn(stem_bn, out=out),
                                       deps=_bn_deps(stem_bn))
                    stem_consts = cache.get(("conv1", "consts"),
                                            lambda out: engine.bound_consts(stem_w, sc, sh, out=out),
                                            deps=(self.conv1.weight,) + _bn_deps(stem_bn)) if scaled else None
                    # a fused projection shortcut contracts the stage input and conv2's output in ONE accumulator:
                    # both are then plain bf16 (exponent 0), so the pooled stem output is stored that way too
                    first_unit = getattr(self, self.res_layers[0])[0]
                    stem_plain = fuse_pool and not train and _fuses_shortcut(first_unit)
                    if fuse_pool:
                        cur = engine.Act(pool.get((cn, hq, wq, 64)), (cn, hq, wq, 64),
                                         torch.bfloat16 if stem_plain else internal, meta.new())
                        ops.append(engine.op_stem(cn, h, w, staged, stem_w, cur, sc, sh, x_meta=staged_meta,
                                                  consts=None if stem_plain else stem_consts,
                                                  scaled_out=scaled and not stem_plain, pool=True))
                        pool.release(staged)
                    else:
                        ops.append(engine.op_stem(cn, h, w, staged, stem_w, stem_out, sc, sh, x_meta=staged_meta,
                                                  consts=stem_consts, scaled_out=scaled, split=split))
                        if not keep_stem:
                            pool.release(staged)
                        # max-pool commutes with the (positive) per-tensor scale: metadata passes through
                        cur = engine.Act(pool.get((cn, hq, wq, 64 * cs)), (cn, hq, wq, 64), internal, stem_out.meta)
                        ops.append(engine.op_maxpool(stem_out, cur, split=split))
                        if keep_stem:
                            stem_rec = dict(staged=staged, stem_out=stem_out, h=h, w=w)  # saved for the stem's backward
                        else:
                            pool.release(stem_out.buf)
                    cur_pooled = True
                else:
                    cur = boundary_act(stages[0] - 1, i0, cn)
                    cur_pooled = False
                for li in stages:
                    lname = self.res_layers[li]
                    stage = getattr(self, lname)
                    pending_z1 = None  # conv1 output of the next block, already produced by a fused bottleneck tail
                    for bi, unit in enumerate(stage):
                        pre = "%s.%d." % (lname, bi)
                        last = bi == len(stage) - 1
                        nb, hb, wb, _ = cur.shape
                        if not train and not split and _fuses_tail(unit):
                            # TDET_OP_BOTTLENECK_TAIL: conv2 -> conv3 + residual + ReLU (-> the next block's conv1) in
                            # one kernel; z2 never reaches memory and the block output is not re-read
                            z1 = pending_z1
                            if z1 is None:
                                z1 = new_act((nb, hb, wb, unit.conv1.out_channels), internal)
                                conv(pre + "conv1", unit.conv1, getattr(unit, unit.norm_names[0]), cur, z1)
                            dst = boundary_act(li, i0, cn) if last else new_act((nb, hb, wb, unit.conv3.out_channels), internal)
                            head = stage[bi + 1] if (FUSE_TAIL_NEXT and unit.conv2.out_channels == 64 and not last and
                                                     _fuses_tail(stage[bi + 1])) else None
                            pending_z1 = tail(pre, unit, z1, cur, dst, head, "%s.%d." % (lname, bi + 1))
                            pool.release(z1.buf)
                            if cur_pooled:
                                pool.release(cur.buf)
                            cur = dst
                            cur_pooled = not last
                            continue
                        hn = engine.conv_out(hb, 3, unit.stride, unit.dilation, unit.dilation)
                        wn = engine.conv_out(wb, 3, unit.stride, unit.dilation, unit.dilation)
                        residual = cur
                        shortcut = None
                        # a bottleneck's projection shortcut runs inside its conv3 launch (second accumulator):
                        # the shortcut tensor is never written or re-read (inference plans)
                        fuse_sc = (not train and not split and cur.dtype == torch.bfloat16 and _fuses_shortcut(unit))
                        if unit.downsample is not None and not fuse_sc:
                            cd = unit.downsample[0].out_channels
                            shortcut = new_act((nb, hn, wn, cd), internal)
                            conv(pre + "downsample", unit.downsample[0], unit.downsample[1], cur, shortcut,
                                 relu=False)
                            residual = shortcut
                        nconv = len(unit.kernel_sizes)
                        src = cur
                        temps = []
                        keep = train and li >= train_from
                        acts = []
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
                                # (the input of a dual-source conv3 is a plain bf16 tensor, see above)
                                dst = new_act((nb, oh, ow, module.out_channels),
                                              torch.bfloat16 if (fuse_sc and ci == nconv - 2) else internal)
                                if not final:
                                    temps.append(dst)
                            if final and fuse_sc:
                                conv(pre + "conv%d" % (ci + 1), module, getattr(unit, unit.norm_names[ci]), src, dst,
                                     shortcut=(pre + "downsample", unit.downsample[0], unit.downsample[1], cur))
                            else:
                                conv(pre + "conv%d" % (ci + 1), module, getattr(unit, unit.norm_names[ci]), src,
                                     dst, residual=residual if final else None)
                            src = dst
                            acts.append(dst)
                        if keep:
                            # saved for backward: block input, every conv output (ReLU masks + wgrad operands)
                            records.append(dict(li=li, bi=bi, pre=pre, unit=unit, xin=cur, acts=acts,
                                                last=last, shortcut=shortcut))
                        else:
                            for t in temps:
                                pool.release(t.buf)
                            if cur_pooled:
                                pool.release(cur.buf)
                        if shortcut is not None and not keep:
                            pool.release(shortcut.buf)  # (kept in training: BatchNorm affine gradients need it)
                        cur = src
                        cur_pooled = not last  # stage outputs live in boundary tensors, never pooled
                    if split and li in self.out_indices:
                        ops.append(engine.op_split_combine(boundary_act(li, 0, n),
                                                           outs_f32[sorted(self.out_indices).index(li)]))
                    if train and li in self.out_indices:
                        # the returned feature map: plain bf16 copy of the internal stage output
                        t, m = boundary[li]
                        ops.append(engine.op_add_mask(cints[li], engine.Act(t, cints[li].shape, torch.bfloat16, m)))
        plan = engine.Plan(ops, [x] + outs + outs_f32, [cache, pool.all_buffers], dev, meta=meta)
        if split:
            return plan, [tuple(o.shape) for o in outs], [tuple(o.shape) for o in outs_f32]
        if train:
            plan.stem_rec = stem_rec
            return plan, [tuple(o.shape) for o in outs], records, cints, geo
        return plan, [tuple(o.shape) for o in outs]

    def _build_plan_gn(self, x, cache):
        """use_gn=True (nn.GroupNorm(32, C) after every conv, layers.py:50-54): GroupNorm needs the statistics of the
        whole conv output, so nothing folds into a GEMM epilogue.  Every conv launches RAW (no affine, no ReLU; fp16
        significands with a device-chosen exponent), TDET_OP_GN_STATS reduces sum / sum of squares per (image,
        group) -- without atomics: bit-reproducible, and independent of the batch an image travels in -- and
        TDET_OP_GN_APPLY normalises, applies gamma / beta, adds the residual and applies the ReLU in one
        pass.  Whole batch per launch, no fusions; inference only."""
        n, _, h_in, w_in = x.shape
        dev = x.device
        tf = getattr(self, "_input_tf", None)
        h, w = h_in, w_in
        tf_scale = tf_shift = None
        if tf is not None:
            means, stds, div = tf
            if div:
                h, w = (h_in + div - 1) // div * div, (w_in + div - 1) // div * div
            tf_scale, tf_shift = cache.get(("input_tf", tf), lambda out: (
                torch.tensor([1.0 / s for s in stds], dtype=torch.float32, device=dev),
                torch.tensor([-m / s for m, s in zip(means, stds)], dtype=torch.float32, device=dev)))
        internal = INTERNAL_DTYPE
        scaled = internal == torch.float16
        ops = []
        pool = _BufferPool(dev)
        norms = [m for m in self.modules() if isinstance(m, nn.GroupNorm)]
        meta = engine.MetaArena(16 + 2 * len(norms), dev)
        # one statistics buffer, reused by every GroupNorm in turn (stream order: its apply has run before the next
        # statistics pass overwrites it)
        stats = torch.empty(engine.gn_stats_numel(n, max(m.num_groups for m in norms)), dtype=torch.float32, device=dev)

        def new_act(shape, dtype):
            return engine.Act(pool.get(shape), shape, dtype, meta.new())

        def raw_conv(name, module, src):
            k = module.kernel_size[0]
            nb, hb, wb, _ = src.shape
            oh = engine.conv_out(hb, k, module.stride[0], module.padding[0], module.dilation[0])
            ow = engine.conv_out(wb, k, module.stride[0], module.padding[0], module.dilation[0])
            if module.groups > 1:
                wgt = cache.get((name, "w", src.dtype), lambda out: engine.pack_grouped_conv_weight(
                    module.weight, module.groups, src.dtype, out=out), deps=(module.weight,))
            else:
                wgt = cache.get((name, "w", src.dtype),
                                lambda out: engine.pack_conv_weight(module.weight, src.dtype, out=out),
                                deps=(module.weight,))
            consts = cache.get((name, "consts_raw", src.dtype), lambda out: engine.bound_consts(wgt, None, None, out=out),
                               deps=(module.weight,)) if scaled else None
            dst = new_act((nb, oh, ow, module.out_channels), internal)
            ops.append(engine.op_conv(src, wgt, dst, k, k, module.stride[0], module.padding[0], module.dilation[0],
                                      relu=False, consts=consts, scaled_out=scaled, groups=module.groups))
            return dst

        def group_norm(name, norm, raw, dst, residual=None, relu=True):
            st = stats[:engine.gn_stats_numel(raw.shape[0], norm.num_groups)]
            gamma, beta = cache.get((name, "gn"), lambda out: _affine_copy(norm, out), deps=(norm.weight, norm.bias))
            ops.append(engine.op_gn_stats(raw, st, norm.num_groups))
            ops.append(engine.op_gn_apply(raw, st, norm.num_groups, gamma, beta, norm.eps, dst, residual=residual,
                                          relu=relu))
            pool.release(raw.buf)

        ho, wo = engine.conv_out(h, 7, 2, 3), engine.conv_out(w, 7, 2, 3)
        hq, wq = engine.conv_out(ho, 3, 2, 1), engine.conv_out(wo, 3, 2, 1)
        staged = pool.get((n,) + engine.stem_staging_dims(ho, wo) + (4,))
        staged_meta = meta.new()
        # fp16 staging and stem weights: a chain of GroupNorms amplifies the stem's operand rounding (bf16 weights
        # alone cost 3e-3 at the outputs of the 64 x 96 fixture); normalised pixels are far inside fp16's range
        ops.append(engine.op_prep(x, staged, ho, wo, y_meta=staged_meta, scale=tf_scale, shift=tf_shift,
                                  padded_hw=(h, w), y_dtype=torch.float16))
        stem_w = cache.get(("conv1", "w", torch.float16),
                           lambda out: engine.pack_stem_weight(self.conv1.weight, out=out, dtype=torch.float16),
                           deps=(self.conv1.weight,))
        stem_consts = cache.get(("conv1", "consts_raw"), lambda out: engine.bound_consts(stem_w, None, None, out=out),
                                deps=(self.conv1.weight,)) if scaled else None
        raw = new_act((n, ho, wo, 64), internal)
        ops.append(engine.op_stem(n, h, w, staged, stem_w, raw, None, None, relu=False, x_meta=staged_meta,
                                  consts=stem_consts, scaled_out=scaled, x_dtype=torch.float16))
        pool.release(staged)
        act = new_act((n, ho, wo, 64), internal)
        group_norm("conv1", getattr(self, self.norm_name), raw, act)
        cur = engine.Act(pool.get((n, hq, wq, 64)), (n, hq, wq, 64), internal, act.meta)  # max-pool keeps the metadata
        ops.append(engine.op_maxpool(act, cur))
        pool.release(act.buf)
        outs = []
        for li, lname in enumerate(self.res_layers):
            stage = getattr(self, lname)
            for bi, unit in enumerate(stage):
                pre = "%s.%d." % (lname, bi)
                last = bi == len(stage) - 1
                residual = cur
                if unit.downsample is not None:
                    raw = raw_conv(pre + "downsample", unit.downsample[0], cur)
                    residual = new_act(raw.shape, internal)
                    group_norm(pre + "downsample", unit.downsample[1], raw, residual, relu=False)
                src = cur
                nconv = len(unit.kernel_sizes)
                for ci in range(nconv):
                    module = getattr(unit, "conv%d" % (ci + 1))
                    raw = raw_conv(pre + "conv%d" % (ci + 1), module, src)
                    final = ci == nconv - 1
                    dst = new_act(raw.shape, internal)
                    group_norm(pre + "conv%d" % (ci + 1), getattr(unit, unit.norm_names[ci]), raw, dst,
                               residual=residual if final else None)
                    if src is not cur:
                        pool.release(src.buf)
                    src = dst
                if residual is not cur:
                    pool.release(residual.buf)
                pool.release(cur.buf)
                cur = src
            if li in self.out_indices:
                # the returned feature map is a bf16 copy: the next stage keeps reading the internal tensor (GroupNorm
                # amplifies storage rounding by |mean| / std of each group, so the residual stream stays in fp16)
                t = engine.nhwc_empty(n, cur.shape[1], cur.shape[2], cur.shape[3], dev)  # re-bound at run time
                outs.append(t)
                ops.append(engine.op_add_mask(cur, engine.Act(t, cur.shape, torch.bfloat16, meta.new())))
        plan = engine.Plan(ops, [x] + outs, [cache, pool.all_buffers, stats], dev, meta=meta)
        return plan, [tuple(o.shape) for o in outs]

    def forward(self, x):
        self._check_supported(x)
        if self.use_gn:
            if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                raise NotImplementedError("training through GroupNorm (use_gn=True) is not on the B200 path: "
                                          "inference only (wrap the call in torch.no_grad() or freeze the module)")
            cache = self._get_operands(x.device)
            key = ("gn", tuple(x.shape), x.dtype, tuple(x.stride()), x.device, INTERNAL_DTYPE,
                   getattr(self, "_input_tf", None))
            entry = self._plans.get(key)
            if entry is None:
                entry = self._build_plan_gn(x, cache)
                self._plans[key] = entry
            plan, out_shapes = entry
            outs = [torch.empty(s, dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
                    for s in out_shapes]
            plan.run([x] + outs)
            self._last_run = (plan, [x] + outs)
            if x.dtype == torch.float32:
                outs = [_upcast(o) for o in outs]
            return outs[0] if len(outs) == 1 else tuple(outs)
        if self.training and torch.is_grad_enabled():
            params = self._trainable_weights()
            if params:
                outs = training.PlanFunction.apply(self, 1, x, *params)
                if x.dtype == torch.float32:
                    outs = [_upcast(o) for o in outs]
                return outs[0] if len(outs) == 1 else tuple(outs)
        cache = self._get_operands(x.device)
        key = (tuple(x.shape), x.dtype, tuple(x.stride()), x.device, INTERNAL_DTYPE, getattr(self, "_input_tf", None))
        entry = self._plans.get(key)
        if entry is None:
            entry = self._build_plan(x, cache)
            self._plans[key] = entry
        if len(entry) == 3:
            # fp32-I/O mode: split-precision stage outputs (kept on the returned tensors for a following B200
            # neck) and their fp32 sums
            plan, split_shapes, f32_shapes = entry
            souts = [torch.empty(s, dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
                     for s in split_shapes]
            outs = [torch.empty(s, dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
                    for s in f32_shapes]
            plan.run([x] + souts + outs)
            self._last_run = (plan, [x] + souts + outs)
            for o, so in zip(outs, souts):
                o._tdet_split = so
            return outs[0] if len(outs) == 1 else tuple(outs)
        plan, out_shapes = entry
        outs = [torch.empty(s, dtype=torch.bfloat16, device=x.device,
                            memory_format=torch.channels_last) for s in out_shapes]
        plan.run([x] + outs)
        self._last_run = (plan, [x] + outs)  # for profiling tools (bench.py)
        if x.dtype == torch.float32:
            outs = [_upcast(o) for o in outs]
        return outs[0] if len(outs) == 1 else tuple(outs)

    # ------------------------------------------------------------------ training path (config 4)
    def set_grad_sync(self, sync):
        """Attach a ``training.BucketAllReduce``: backward then all-reduces one flat fp32 bucket per
        stage (deepest first) as soon as that stage's wgrad kernels are enqueued."""
        self._grad_sync = sync
        if sync is not None:
            sync.attach(self, next(self.parameters()).device)

    def _trainable_weights(self):
        """Parameters that get gradients, in module order: the conv weights of the trainable stages and --
        with bn_frozen=False -- the affine parameters of their (eval-mode) BatchNorms.  Supported: a frozen
        stem (frozen_stages >= 0), a frozen prefix of stages and a fully trainable suffix; per stage the BN
        affine parameters are either all trainable or all frozen."""
        stem_norm = getattr(self, self.norm_name)
        first = None
        params = []
        bn_train = []
        for li, lname in enumerate(self.res_layers):
            stage = getattr(self, lname)
            convs = [m for m in stage.modules() if isinstance(m, nn.Conv2d)]
            bns = [m for m in stage.modules() if isinstance(m, nn.BatchNorm2d)]
            flags = [m.weight.requires_grad for m in convs]
            bflags = [p.requires_grad for m in bns for p in (m.weight, m.bias)]
            if any(flags):
                if not all(flags):
                    raise NotImplementedError("partially frozen stage %s" % lname)
                if any(bflags) and not all(bflags):
                    raise NotImplementedError("partially frozen BatchNorm parameters in stage %s" % lname)
                if first is None:
                    first = li
                params.extend(m.weight for m in convs)
                if all(bflags) and bflags:
                    for m in bns:
                        params.extend((m.weight, m.bias))
                bn_train.append(bool(bflags) and all(bflags))
            else:
                if first is not None:
                    raise NotImplementedError("frozen stage %s after a trainable one" % lname)
                bn_train.append(False)
        # the stem (frozen_stages = -1): 7x7 wgrad + max-pool/ReLU backward, only below a trainable stage 1
        stem_bn_flags = [stem_norm.weight.requires_grad, stem_norm.bias.requires_grad]
        self._train_stem = bool(self.conv1.weight.requires_grad)
        self._train_stem_bn = all(stem_bn_flags)
        if self._train_stem or any(stem_bn_flags):
            if first != 0:
                raise NotImplementedError("a trainable stem below a frozen stage 1")
            if not self._train_stem or (any(stem_bn_flags) and not all(stem_bn_flags)):
                raise NotImplementedError("partially frozen stem (conv1 / bn1)")
            params.append(self.conv1.weight)
            if self._train_stem_bn:
                params.extend((stem_norm.weight, stem_norm.bias))
        self._train_from = first
        self._train_bn = tuple(bn_train)
        return params

    def saved_activations(self):
        """{name: fp32 NCHW CPU tensor} of what the last training forward saved for backward (block
        inputs ``layerL.B.in`` and every conv's stored output ``layerL.B.convK``).  Test/debug API:
        the gradient tests feed these to their fp32 checker so that ReLU masks match the kernels'."""
        plan, out_shapes, records, cints, geo = self._train_state["entry"]
        raw = plan.meta.tensor.cpu()
        base = plan.meta.tensor.data_ptr()

        def fetch(act):
            n, h, w, c = act.shape
            flat = act.buf[act.offset:act.offset + n * h * w * c].view(act.dtype)
            e = int(raw[(act.meta - base) // 8, 0]) if act.meta is not None else 0
            return flat.view(n, h, w, c).permute(0, 3, 1, 2).float().cpu() * (2.0 ** e)

        saved = {}
        if getattr(plan, "stem_rec", None) is not None:
            saved["stem.out"] = fetch(plan.stem_rec["stem_out"])
        for r in records:
            saved[r["pre"] + "in"] = fetch(r["xin"])
            for ci, a in enumerate(r["acts"]):
                saved[r["pre"] + "conv%d" % (ci + 1)] = fetch(a)
        return saved

    def _train_forward(self, inputs, params):
        (x,) = inputs
        cache = self._get_operands(x.device)
        key = ("train", tuple(x.shape), x.dtype, tuple(x.stride()), x.device, self._train_from, self._train_bn,
               self._train_stem, self._train_stem_bn, getattr(self, "_input_tf", None))
        entry = self._plans.get(key)
        if entry is None:
            entry = self._build_plan(x, cache, train_from=self._train_from)
            self._plans[key] = entry
        plan, out_shapes, records, boundary, geo = entry
        outs = [torch.empty(s, dtype=torch.bfloat16, device=x.device,
                            memory_format=torch.channels_last) for s in out_shapes]
        plan.run([x] + outs)
        self._last_run = (plan, [x] + outs)
        self._train_serial = getattr(self, "_train_serial", 0) + 1
        state = dict(key=key, entry=entry, outs=outs, params=list(params), serial=self._train_serial, x=x)
        self._train_state = state
        return outs, state

    def _build_bwd_plan(self, state, cache):
        """Backward plan for the saved forward `state`: stages deepest first, blocks last to first."""
        plan, out_shapes, records, cints, geo = state["entry"]
        dev = state["x"].device
        outs = state["outs"]
        scaled = INTERNAL_DTYPE == torch.float16
        # gradient tensors are stored in the format of the saved activations (one format per MMA): fp16
        # significands with a per-tensor exponent chosen on the device (8x finer than bf16), or plain bf16
        n_blocks = sum(len(getattr(self, l)) for l in self.res_layers)
        meta = engine.MetaArena(16 + 8 * n_blocks, dev) if scaled else None
        bb = training.BackwardBuilder(dev, cache, INTERNAL_DTYPE, meta)
        out_pos = {li: i for i, li in enumerate(sorted(self.out_indices))}
        g_ext = [engine.nhwc_empty(*_nhwc_shape(o), dev) for o in outs]  # placeholders, re-bound per run
        ext_acts = {}

        def ext_grad(li):
            """Gradient autograd hands in for returned stage li (plain bf16); its |max| is measured once so
            that the consuming kernel can bound its own output."""
            if li not in ext_acts:
                m = meta.new() if scaled else None
                a = engine.Act(g_ext[out_pos[li]], cints[li].shape, torch.bfloat16, m)
                if scaled:
                    bb.ops.append(engine.op_amax(engine.Act(a.buf, a.shape, a.dtype, None), m))
                ext_acts[li] = a
            return ext_acts[li]

        def stage_act(li):
            return cints[li]

        def scale_of(name, bn):
            return cache.get((name, "bn"), lambda out: engine.fold_bn(bn, out=out), deps=_bn_deps(bn))[0]

        def live(act):
            return act

        nst = len(self.res_layers)
        first = self._train_from
        buckets = []
        segments = []  # (first op, last op, bucket index)
        by_stage = {}
        for r in records:
            by_stage.setdefault(r["li"], []).append(r)
        if (nst - 1) not in out_pos:
            raise NotImplementedError("training needs the last stage among out_indices")
        g_cur = None  # masked gradient w.r.t. the output of the block being processed
        for li in range(nst - 1, first - 1, -1):
            seg_start = len(bb.ops)
            stage = getattr(self, self.res_layers[li])
            convs = [m for m in stage.modules() if isinstance(m, nn.Conv2d)]
            bn_on = self._train_bn[li]
            bparams = [m.weight for m in convs]
            if bn_on:
                for m in stage.modules():
                    if isinstance(m, nn.BatchNorm2d):
                        bparams.extend((m.weight, m.bias))  # adjacent: {dgamma[C], dbeta[C]} is one accumulator
            bucket = training.GradBucket(bparams, dev)
            buckets.append(bucket)

            def bn_grad(name, bn, g, a, b=None, bucket=bucket):
                """dgamma / dbeta of an eval-mode BatchNorm with trainable affine parameters."""
                i = bucket.index_of(bn.weight)
                c = bn.weight.numel()
                dw = bucket.flat[bucket.offsets[i]:bucket.offsets[i] + 2 * c]
                assert bucket.offsets[i + 1] == bucket.offsets[i] + c
                gamma, beta = cache.get((name, "affine"), lambda out: _affine_copy(bn, out), deps=(bn.weight, bn.bias))
                bb.ops.append(engine.op_bn_affine_grad(g, a, gamma, beta, dw, b=b))
            blocks = by_stage[li]
            if li == nst - 1:
                # gradient of the top stage output arrives from autograd only: apply its ReLU mask
                top = stage_act(li)
                g_cur = bb.new_act(top.shape)
                bb.ops.append(engine.op_add_mask(ext_grad(li), g_cur, mask=top, scaled_out=scaled))
            for r in reversed(blocks):
                unit, pre = r["unit"], r["pre"]
                xin, acts = live(r["xin"]), [live(a) for a in r["acts"]]
                nconv = len(unit.kernel_sizes)
                gM = g_cur
                g = gM
                for ci in range(nconv - 1, -1, -1):
                    module = getattr(unit, "conv%d" % (ci + 1))
                    name = pre + "conv%d" % (ci + 1)
                    sc = scale_of(name, getattr(unit, unit.norm_names[ci]))
                    x_ci = xin if ci == 0 else acts[ci - 1]
                    bb.wgrad(name, module, sc, x_ci, g, bucket.view(bucket.index_of(module.weight)))
                    if bn_on:
                        if ci == nconv - 1:
                            # y = out - residual operand wherever the (masked) gradient is non-zero
                            res_op = live(r["shortcut"]) if unit.downsample is not None else xin
                            bn_grad(name, getattr(unit, unit.norm_names[ci]), g, acts[ci], b=res_op)
                        else:
                            bn_grad(name, getattr(unit, unit.norm_names[ci]), g, acts[ci])
                    if ci > 0:
                        g_next = bb.dgrad(name, module, sc, g, acts[ci - 1].shape, mask=acts[ci - 1],
                                          deps=_bn_deps(getattr(unit, unit.norm_names[ci])))
                        if g is not gM:
                            bb.release(g)
                        g = g_next
                ds = unit.downsample
                if ds is not None:
                    name = pre + "downsample"
                    sc_ds = scale_of(name, ds[1])
                    bb.wgrad(name, ds[0], sc_ds, xin, gM, bucket.view(bucket.index_of(ds[0].weight)))
                    if bn_on:
                        bn_grad(name, ds[1], gM, live(r["shortcut"]))
                # gradient w.r.t. the block input: needed unless the producer is frozen
                is_first_block = r["bi"] == 0
                need_gin = not (is_first_block and li == first) or (li == 0 and self._train_stem)
                g_in = None
                if need_gin:
                    module = unit.conv1
                    name = pre + "conv1"
                    sc = scale_of(name, getattr(unit, unit.norm_names[0]))
                    residual, coarse, gd = None, None, None
                    if ds is not None:
                        name_ds = pre + "downsample"
                        wd = bb.dgrad_weight(name_ds, ds[0], scale_of(name_ds, ds[1]), deps=_bn_deps(ds[1]))
                        s_ds = ds[0].stride[0]
                        if s_ds == 1:
                            gd = bb.new_act(xin.shape)
                        elif s_ds == 2:
                            gd = bb.new_act((xin.shape[0], gM.shape[1], gM.shape[2], xin.shape[3]))
                        else:
                            raise NotImplementedError("shortcut stride %d" % s_ds)
                        bb.conv_dgrad_op(name_ds, ds[0], wd, gM, gd, 1, 0, 1, deps=_bn_deps(ds[1]))
                        if s_ds == 1:
                            residual = gd
                        else:
                            coarse = gd
                    else:
                        residual = gM
                    merged = None
                    if is_first_block and (li - 1) in out_pos:
                        # the stage input is a returned feature map: add the gradient autograd hands in
                        ext = ext_grad(li - 1)
                        if residual is None:
                            residual = ext
                        else:
                            merged = bb.new_act(xin.shape)
                            bb.ops.append(engine.op_add_mask(residual, merged, residual=ext, scaled_out=scaled))
                            residual = merged
                    g_in = bb.dgrad(name, module, sc, g, xin.shape, residual=residual, coarse=coarse, mask=xin,
                                    deps=_bn_deps(getattr(unit, unit.norm_names[0])))
                    bb.release(gd)
                    bb.release(merged)
                if g is not gM:
                    bb.release(g)
                bb.release(gM)
                g_cur = g_in
            segments.append([seg_start, len(bb.ops), len(buckets) - 1])
        if self._train_stem:
            # stem: max-pool + ReLU backward (scatter to the window maxima), bn1 affine gradients, 7x7 wgrad
            seg_start = len(bb.ops)
            rec = plan.stem_rec
            stem_bn = getattr(self, self.norm_name)
            sparams = [self.conv1.weight] + ([stem_bn.weight, stem_bn.bias] if self._train_stem_bn else [])
            bucket = training.GradBucket(sparams, dev)
            buckets.append(bucket)
            s_out = rec["stem_out"]
            d_s = engine.Act(torch.empty(s_out.shape[0] * s_out.shape[1] * s_out.shape[2] * 64, dtype=torch.bfloat16,
                                         device=dev), s_out.shape, torch.bfloat16)
            bb.buffers.append(d_s.buf)
            bb.ops.append(engine.op_maxpool_bwd(s_out, g_cur, d_s))
            if self._train_stem_bn:
                gamma, beta = cache.get(("conv1", "affine"), lambda out: _affine_copy(stem_bn, out),
                                        deps=(stem_bn.weight, stem_bn.bias))
                c = stem_bn.weight.numel()
                dwb = bucket.flat[bucket.offsets[1]:bucket.offsets[1] + 2 * c]
                bb.ops.append(engine.op_bn_affine_grad(d_s, s_out, gamma, beta, dwb))
            bb.ops.append(engine.op_stem_wgrad(s_out.shape[0], rec["h"], rec["w"], rec["staged"], d_s, bucket.view(0),
                                               scale=scale_of("conv1", stem_bn)))
            segments.append([seg_start, len(bb.ops), len(buckets) - 1])
        ops, shift = bb.finalize()
        # every bucket is re-zeroed at the start of the run (1x1 wgrads accumulate straight into it)
        head = [engine.op_zero(b.flat) for b in buckets]
        ops = head + ops
        shift += len(head)
        segments = [(a + shift, b + shift, k) for a, b, k in segments]
        segments[0] = (0, segments[0][1], segments[0][2])
        bplan = engine.Plan(ops, g_ext, [cache, bb.buffers, bb.acc_ws, [b.flat for b in buckets]], dev, meta=meta)
        return bplan, buckets, segments

    def _train_backward(self, state, gouts):
        if state["serial"] != getattr(self, "_train_serial", 0):
            raise RuntimeError("ResNet backward after a newer training forward of the same module: the "
                               "saved activations live in the plan's static buffers (one forward per backward)")
        cache = self._get_operands(state["x"].device)
        bkey = ("bwd",) + state["key"]
        entry = self._plans.get(bkey, group=state["key"])
        if entry is None:
            entry = self._plans.put(bkey, self._build_bwd_plan(state, cache), group=state["key"])
        bplan, buckets, segments = entry
        outs = state["outs"]
        ext = [training.as_grad_nhwc(g, o) for g, o in zip(gouts, outs)]
        sync = getattr(self, "_grad_sync", None)
        reduced = []
        for a, b, k in segments:
            if sync is not None:
                sync.guard(buckets[k].flat)
            bplan.run_range(ext, a, b)
            # the plan's accumulators are reused by the next backward: autograd gets a copy, which is
            # made (and all-reduced) on the side stream while the next segment computes
            reduced.append(sync.reduce(buckets[k].flat, buckets[k].params) if sync is not None else buckets[k].flat.clone())
        if sync is not None:
            sync.module_done()
        self._last_bwd_run = (bplan, ext)
        lookup = {}
        for (a, b, k), flat in zip(segments, reduced):
            bucket = buckets[k]
            for i, p in enumerate(bucket.params):
                lookup[id(p)] = flat[bucket.offsets[i]:bucket.offsets[i] + p.numel()].view(p.shape)
        grads = [lookup[id(p)] for p in state["params"]]
        return [None], grads


def _nhwc_shape(t):
    n, c, h, w = t.shape
    return n, h, w, c


# Storage format of the backbone's internal activations: float16 = fp16 significand + per-tensor
# power-of-two exponent (default; ~8x lower rounding error than bf16 at the same tensor-core rate),
# bfloat16 = plain bf16 everywhere (TDET_INTERNAL_DTYPE=bf16).
INTERNAL_DTYPE = torch.bfloat16 if os.environ.get("TDET_INTERNAL_DTYPE", "fp16").lower() in (
    "bf16", "bfloat16") else torch.float16

# fp32 inputs in eval mode run the split-precision path (fp32-I/O tolerance 1e-4, ~3x the MMA work and 2x the
# bytes); TDET_FP32_IO=bf16 keeps fp32 tensors at the module boundary only (bf16-accurate, fast).
FP32_IO_SPLIT = os.environ.get("TDET_FP32_IO", "split").lower() == "split"
# the stem kernel applies the 3x3/2 max-pool in its epilogue (TDET_FLAG_POOL); 0 = separate TDET_OP_MAXPOOL launch
FUSE_STEM_POOL = os.environ.get("TDET_STEM_POOL", "1") != "0"

# a stage's projection shortcut is contracted inside the first bottleneck's conv3 launch (TDET_FLAG_DUAL)
FUSE_SHORTCUT = os.environ.get("TDET_FUSE_SHORTCUT", "1") != "0"

# conv2 -> conv3 + residual -> next conv1 of the identity-shortcut bottlenecks of layer1 as one kernel
FUSE_TAIL = os.environ.get("TDET_FUSE_TAIL", "1") != "0"
# ... and the next block's conv1 in the same kernel (measured: no faster than the tail-only kernel + a separate conv1,
# the tail-only variant keeps W2 resident; kept as an experiment switch)
FUSE_TAIL_NEXT = os.environ.get("TDET_FUSE_TAIL_NEXT", "0") != "0"
# the planes = 128 tail kernel for layer2's identity blocks (bottleneck_tail2.cuh).  Off by default: same-box A/B at
# batch 16 @800x1344 measured the three fused blocks 50 us EACH slower than conv2 + conv3 launched separately
# (265 vs 88 + 127 us in situ; 2785 vs 2855 img/s for the whole step).  The kernel and its tests stay (TDET_FUSE_TAIL2=1).
FUSE_TAIL2 = os.environ.get("TDET_FUSE_TAIL2", "0") != "0"

# Images per chunk for stage 1, 2, ... ("0" or missing = whole batch); TDET_CHUNKS overrides.
DEFAULT_CHUNKS = "0"


def _fuses_shortcut(unit):
    """True if the unit's projection shortcut can run inside its last conv's launch (TDET_FLAG_DUAL): a bottleneck
    whose last conv is a dense 1x1 with a multiple of 256 output channels and whose shortcut is a 1x1 conv."""
    if not FUSE_SHORTCUT or unit.downsample is None or len(unit.kernel_sizes) < 2 or unit.kernel_sizes[-1] != 1:
        return False
    last = getattr(unit, "conv%d" % len(unit.kernel_sizes))
    ds = unit.downsample[0]
    return (last.out_channels % 256 == 0 and last.groups == 1 and last.stride[0] == 1 and
            ds.kernel_size[0] == 1 and ds.padding[0] == 0 and ds.groups == 1 and ds.in_channels % 64 == 0)


def _fuses_tail(unit):
    """True if conv2 -> conv3 + residual (-> the next block's conv1) of this unit run as ONE kernel
    (TDET_OP_BOTTLENECK_TAIL): an identity-shortcut bottleneck with planes = 64 (layer1 of ResNet-50/101/152) or
    planes = 128 (layer2; tail only)."""
    if not FUSE_TAIL or unit.downsample is not None or tuple(unit.kernel_sizes) != (1, 3, 1):
        return False
    c1, c2, c3 = unit.conv1, unit.conv2, unit.conv3
    pl = c2.out_channels
    if pl != 64 and not (pl == 128 and FUSE_TAIL2):
        return False
    return (c1.in_channels == 4 * pl and c1.out_channels == pl and c1.stride[0] == 1 and c1.groups == 1 and
            c2.in_channels == pl and c2.stride[0] == 1 and c2.padding[0] == 1 and
            c2.dilation[0] == 1 and c2.groups == 1 and c3.in_channels == pl and c3.out_channels == 4 * pl and
            c3.groups == 1)


def _sum_into(a, b, out):
    if out is None:
        return (a + b).contiguous()
    torch.add(a, b, out=out)
    return out


def _affine_copy(bn, out):
    """fp32 device copies of an eval-mode BatchNorm's (gamma, beta), refreshed in place."""
    g, b = bn.weight.detach().float(), bn.bias.detach().float()
    if out is None:
        return g.contiguous().clone(), b.contiguous().clone()
    out[0].copy_(g)
    out[1].copy_(b)
    return out


def _bn_deps(bn):
    return (bn.weight, bn.bias, bn.running_mean, bn.running_var)


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
