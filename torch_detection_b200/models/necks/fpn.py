"""Feature Pyramid Network neck, B200 execution -- drop-in for reference ``models/necks/fpn.py``.

Constructor arguments, attributes (``lateral_convs`` / ``fpn_convs`` ``ModuleList``s of
``ConvModule`` with a ``.conv``), ``state_dict`` keys, ``init_weights`` and the returned tuple are
those of the reference ``FPN`` (fpn.py:8-125).  ``forward`` compiles, once per input geometry, the
op list

    lateral[j]  = conv1x1(C[j]) + bias (+ nearest-x2 upsample of lateral[j+1])   coarse -> fine
    P[j]        = conv3x3(lateral[j]) + bias
    extra       = stride-2 subsample of the last P (fpn.py:114-116) or stride-2 3x3 convs (RetinaNet)

where the top-down add (fpn.py:98-101) is fused into the lateral conv's epilogue in fp32, so each
merged lateral is written exactly once.  Like the reference, mismatching pyramid levels
(fine != 2 x coarse, e.g. a raw 800x1333 image) raise ``RuntimeError``; nothing is cropped silently.
"""
import torch
import torch.nn as nn

from ... import engine, training
from ...registry import NECKS
from ..utils import ConvModule, xavier_init, constant_init


@NECKS.register_module
class FPN(nn.Module):

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1,
                 add_extra_convs=False, normalize=None, use_gn=False):
        super(FPN, self).__init__()
        assert isinstance(in_channels, list)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_ins = len(in_channels)
        self.num_outs = num_outs
        self.with_bias = normalize is None
        if end_level == -1:
            self.backbone_end_level = self.num_ins
            assert num_outs >= self.num_ins - start_level
        else:
            # if end_level < inputs, no extra level is allowed
            self.backbone_end_level = end_level
            assert end_level <= len(in_channels)
            assert num_outs == end_level - start_level
        self.start_level = start_level
        self.end_level = end_level
        self.add_extra_convs = add_extra_convs

        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        # lateral and output conv of a level are created pairwise (fpn.py:43-61): RNG order matters
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(ConvModule(in_channels[i], out_channels, 1,
                                                 normalize=normalize, bias=self.with_bias,
                                                 use_gn=use_gn))
            self.fpn_convs.append(ConvModule(out_channels, out_channels, 3, padding=1,
                                             normalize=normalize, bias=self.with_bias,
                                             use_gn=use_gn))
        extra_levels = num_outs - self.backbone_end_level + self.start_level
        if add_extra_convs and extra_levels >= 1:
            for i in range(extra_levels):
                cin = in_channels[self.backbone_end_level - 1] if i == 0 else out_channels
                self.fpn_convs.append(ConvModule(cin, out_channels, 3, stride=2, padding=1,
                                                 normalize=normalize, bias=self.with_bias,
                                                 use_gn=use_gn))
        self._plans = engine.PlanCache()
        self._operands = None
        self._operand_key = None

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                xavier_init(m, distribution="uniform")
            if isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                constant_init(m, 1)

    # ------------------------------------------------------------------ B200 execution
    def _get_operands(self, device):
        """Packed bf16 weights and fp32 biases; re-derived in place when a parameter changes
        (engine.OperandCache), so plans and their TMA descriptors survive optimizer steps."""
        if self._operands is not None and self._operand_key == device:
            force = getattr(self, "_operands_stale", False) or (engine.REFRESH_EVERY_STEP and self.training)
            self._operands.refresh(force=force)
            self._operands_stale = False
            return self._operands
        cache = engine.OperandCache()
        for kind, convs in self._conv_groups():
            for j, cm in enumerate(convs):
                conv = cm.conv
                cache.get("%s%d.w" % (kind, j),
                          lambda out, conv=conv: engine.pack_conv_weight(conv.weight, out=out),
                          deps=(conv.weight,))
                if cm.with_norm and isinstance(cm.norm, nn.GroupNorm):
                    # GroupNorm runs after the raw conv (TDET_OP_GN_STATS / _APPLY): its operands are made by
                    # _build_plan_gn
                    pass
                elif cm.with_norm:
                    # eval-mode BatchNorm folded to the epilogue's scale / shift (a conv bias folds into shift)
                    norm = cm.norm
                    cache.get("%s%d.bn" % (kind, j),
                              lambda out, conv=conv, norm=norm: _fold_conv_bn(conv, norm, out),
                              deps=tuple(t for t in (norm.weight, norm.bias, norm.running_mean, norm.running_var,
                                                     conv.bias) if t is not None))
                elif conv.bias is not None:
                    cache.get("%s%d.b" % (kind, j),
                              lambda out, conv=conv: _bias_copy(conv.bias, out), deps=(conv.bias,))
        self._operands = cache
        self._operand_key = device
        self._plans = engine.PlanCache()
        return cache

    def invalidate_operands(self):
        """See ResNet.invalidate_operands: call after in-place parameter updates made through ``.data``."""
        self._operands_stale = True

    def _conv_groups(self):
        return (("lat", self.lateral_convs), ("out", self.fpn_convs))

    def _build_plan(self, feats, operands, split=False):
        """split: fp32-I/O mode -- `feats` are split-precision tensors (bf16 hi|lo pairs, 2x channels), every conv
        runs the 3-pass hi/lo GEMM and the returned tensors are the fp32 sums of split outputs."""
        dev = feats[0].device
        n = feats[0].shape[0]
        cs = 2 if split else 1
        used = feats[self.start_level:self.backbone_end_level]
        nl = len(used)
        shapes = [(t.shape[0], t.shape[2], t.shape[3], t.shape[1] // cs) for t in used]  # n,h,w,c (logical)

        def wkey(key):
            if not split:
                return operands.value(key + ".w")
            kind, j = ("lat", int(key[3:])) if key.startswith("lat") else ("out", int(key[3:]))
            conv = (self.lateral_convs if kind == "lat" else self.fpn_convs)[j].conv
            return operands.get(key + ".wsplit", lambda out: engine.pack_conv_weight_split(conv.weight, out=out),
                                deps=(conv.weight,))
        for j in range(nl - 1, 0, -1):
            for dim, a, b in ((2, shapes[j - 1][1], 2 * shapes[j][1]), (3, shapes[j - 1][2], 2 * shapes[j][2])):
                if a != b:
                    # same failure class/message as the reference's in-place add (fpn.py:100-101)
                    raise RuntimeError(
                        "The size of tensor a (%d) must match the size of tensor b (%d) at "
                        "non-singleton dimension %d" % (a, b, dim))
        co = self.out_channels
        ops = []
        srcs = [engine.Act(t, shapes[j], torch.bfloat16) for j, t in enumerate(used)]
        lats = [None] * nl
        for j in range(nl - 1, -1, -1):
            nb, h, w, c = shapes[j]
            lats[j] = engine.Act(torch.empty(nb * h * w * co * cs, dtype=torch.bfloat16, device=dev),
                                 (nb, h, w, co), torch.bfloat16)
            ops.append(engine.op_conv(srcs[j], wkey("lat%d" % j), lats[j], 1, 1, 1, 0, 1,
                                      **_epi(operands, "lat%d" % j),
                                      coarse=lats[j + 1] if j < nl - 1 else None, split=split))
        if split:
            outs, keep = FPN._emit_outputs(self, ops, operands, lats, shapes, dev, wkey=wkey, split=True)
        else:
            outs, keep = self._emit_outputs(ops, operands, lats, shapes, dev)
        if self.num_outs > nl:
            if not self.add_extra_convs:
                for _ in range(self.num_outs - nl):
                    nb, h, w, _ = outs[-1].shape
                    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
                    o = engine.Act(engine.nhwc_empty(nb, oh, ow, co * cs, dev), (nb, oh, ow, co), torch.bfloat16)
                    ops.append(engine.op_subsample(outs[-1], o, split=split))
                    outs.append(o)
            else:
                top = feats[self.backbone_end_level - 1]
                src = engine.Act(top, (top.shape[0], top.shape[2], top.shape[3], top.shape[1] // cs), torch.bfloat16)
                for j in range(nl, self.num_outs):
                    nb, h, w, c = src.shape
                    oh, ow = engine.conv_out(h, 3, 2, 1), engine.conv_out(w, 3, 2, 1)
                    o = engine.Act(engine.nhwc_empty(nb, oh, ow, co * cs, dev), (nb, oh, ow, co), torch.bfloat16)
                    # the reference applies ReLU *in place* to P_j before the next extra conv
                    # (fpn.py:123-124), so every extra level that feeds another one is returned
                    # post-ReLU: fold that ReLU into the producing conv's epilogue.
                    ops.append(engine.op_conv(src, wkey("out%d" % j), o, 3, 3, 2, 1, 1,
                                              **_epi(operands, "out%d" % j),
                                              relu=(j < self.num_outs - 1), split=split))
                    outs.append(o)
                    src = o
        f32 = []
        if split:
            for o in outs:
                nb, h, w, c = o.shape
                t = engine.nhwc_empty(nb, h, w, c, dev, torch.float32)
                ops.append(engine.op_split_combine(o, t))
                f32.append(t)
        ext = list(feats) + [o.buf for o in outs] + f32
        plan = engine.Plan(ops, ext, [operands, [l.buf for l in lats], keep], dev)
        plan.lats = lats  # merged laterals: the saved activations of the training path
        if split:
            return plan, [tuple(o.buf.shape) for o in outs], [tuple(t.shape) for t in f32]
        return plan, [tuple(o.buf.shape) for o in outs]

    def _uses_gn(self):
        return any(isinstance(m, nn.GroupNorm) for m in self.modules())

    def _build_plan_gn(self, feats, operands):
        """normalize=..., use_gn=True: every ConvModule is conv -> GroupNorm(32, C) (layers.py:120-124).  GroupNorm
        needs the statistics of the whole conv output, so each conv launches raw (fp16 significands with a
        device-chosen exponent; the inputs' max |x| comes from one TDET_OP_AMAX pass), TDET_OP_GN_STATS reduces the
        per-(image, group) sums and TDET_OP_GN_APPLY normalises and -- for the laterals -- adds the nearest-upsampled
        coarser level in the same pass (fpn.py:98-101)."""
        if type(self) is not FPN:
            raise NotImplementedError("%s with GroupNorm is not on the B200 path" % type(self).__name__)
        dev = feats[0].device
        n = feats[0].shape[0]
        used = feats[self.start_level:self.backbone_end_level]
        nl = len(used)
        shapes = [(t.shape[0], t.shape[2], t.shape[3], t.shape[1]) for t in used]
        for j in range(nl - 1, 0, -1):
            for dim, a, b in ((2, shapes[j - 1][1], 2 * shapes[j][1]), (3, shapes[j - 1][2], 2 * shapes[j][2])):
                if a != b:
                    raise RuntimeError(
                        "The size of tensor a (%d) must match the size of tensor b (%d) at "
                        "non-singleton dimension %d" % (a, b, dim))
        co = self.out_channels
        n_norm = len(self.lateral_convs) + len(self.fpn_convs)
        meta = engine.MetaArena(nl + 3 * n_norm + 4, dev)
        groups = self.lateral_convs[0].norm.num_groups
        stats = torch.empty(engine.gn_stats_numel(n, groups), dtype=torch.float32, device=dev)  # reused in stream order
        ops = []
        keep = [stats]

        def conv_gn(key, cm, src, dst, stride=1, coarse=None, relu=False):
            """dst = act(GroupNorm(conv(src) + bias) + up2(coarse))"""
            conv, norm = cm.conv, cm.norm
            k = conv.kernel_size[0]
            nb, h, w, _ = src.shape
            oh, ow = engine.conv_out(h, k, stride, conv.padding[0]), engine.conv_out(w, k, stride, conv.padding[0])
            wgt = operands.get((key, "w", src.dtype),
                               lambda out: engine.pack_conv_weight(conv.weight, src.dtype, out=out), deps=(conv.weight,))
            bias = operands.get((key, "b"), lambda out: _bias_copy(conv.bias, out),
                                deps=(conv.bias,)) if conv.bias is not None else None
            deps = (conv.weight,) + ((conv.bias,) if conv.bias is not None else ())
            consts = operands.get((key, "consts_raw", src.dtype),
                                  lambda out: engine.bound_consts(wgt, None, bias, out=out), deps=deps)
            raw = engine.Act(torch.empty(nb * oh * ow * co, dtype=torch.bfloat16, device=dev), (nb, oh, ow, co),
                             torch.float16, meta.new())
            keep.append(raw.buf)
            ops.append(engine.op_conv(src, wgt, raw, k, k, stride, conv.padding[0], 1, shift=bias, consts=consts,
                                      scaled_out=True))
            st = stats[:engine.gn_stats_numel(nb, norm.num_groups)]
            gamma, beta = operands.get((key, "gn"), lambda out: _gn_affine(norm, out), deps=(norm.weight, norm.bias))
            ops.append(engine.op_gn_stats(raw, st, norm.num_groups))
            ops.append(engine.op_gn_apply(raw, st, norm.num_groups, gamma, beta, norm.eps, dst, coarse=coarse, relu=relu))

        srcs = []
        for j, t in enumerate(used):
            a = engine.Act(t, shapes[j], torch.bfloat16, meta.new())
            ops.append(engine.op_amax(engine.Act(t, shapes[j], torch.bfloat16), a.meta))
            srcs.append(a)
        lats = [None] * nl
        for j in range(nl - 1, -1, -1):
            nb, h, w, c = shapes[j]
            lats[j] = engine.Act(torch.empty(nb * h * w * co, dtype=torch.bfloat16, device=dev), (nb, h, w, co),
                                 torch.float16, meta.new())
            conv_gn("lat%d" % j, self.lateral_convs[j], srcs[j], lats[j], coarse=lats[j + 1] if j < nl - 1 else None)
        outs = []
        for j in range(nl):
            nb, h, w, _ = shapes[j]
            o = engine.Act(engine.nhwc_empty(nb, h, w, co, dev), (nb, h, w, co), torch.bfloat16, meta.new())
            conv_gn("out%d" % j, self.fpn_convs[j], lats[j], o)
            outs.append(o)
        if self.num_outs > nl:
            if not self.add_extra_convs:
                for _ in range(self.num_outs - nl):
                    nb, h, w, _ = outs[-1].shape
                    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
                    o = engine.Act(engine.nhwc_empty(nb, oh, ow, co, dev), (nb, oh, ow, co), torch.bfloat16)
                    ops.append(engine.op_subsample(outs[-1], o))
                    outs.append(o)
            else:
                src = srcs[-1]
                for j in range(nl, self.num_outs):
                    nb, h, w, c = src.shape
                    oh, ow = engine.conv_out(h, 3, 2, 1), engine.conv_out(w, 3, 2, 1)
                    o = engine.Act(engine.nhwc_empty(nb, oh, ow, co, dev), (nb, oh, ow, co), torch.bfloat16, meta.new())
                    # (every extra level that feeds another one is returned post-ReLU, see _build_plan)
                    conv_gn("out%d" % j, self.fpn_convs[j], src, o, stride=2, relu=(j < self.num_outs - 1))
                    outs.append(o)
                    src = o
        ext = list(feats) + [o.buf for o in outs]
        plan = engine.Plan(ops, ext, [operands, [l.buf for l in lats], keep], dev, meta=meta)
        plan.lats = lats
        return plan, [tuple(o.buf.shape) for o in outs]

    def _emit_outputs(self, ops, operands, lats, shapes, dev, wkey=None, split=False):
        """P_j = conv3x3(merged lateral j) + bias (fpn.py:106-108).  Returns (output Acts, buffers to keep
        alive); subclasses extend the pyramid here."""
        co = self.out_channels
        cs = 2 if split else 1
        outs = []
        for j in range(len(lats)):
            nb, h, w, _ = shapes[j]
            o = engine.Act(engine.nhwc_empty(nb, h, w, co * cs, dev), (nb, h, w, co), torch.bfloat16)
            wgt = wkey("out%d" % j) if wkey is not None else operands.value("out%d.w" % j)
            ops.append(engine.op_conv(lats[j], wgt, o, 3, 3, 1, 1, 1, **_epi(operands, "out%d" % j), split=split))
            outs.append(o)
        return outs, []

    @staticmethod
    def _as_bf16_nhwc(t):
        engine.require_cuda(t, "FPN input")
        src = getattr(t, "_tdet_bf16", None)
        if src is not None:
            return src
        if t.dtype == torch.bfloat16 and t.is_contiguous(memory_format=torch.channels_last):
            return t
        # foreign producer (NCHW-contiguous and/or fp32): one layout/precision adaptation copy
        return t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        if self.training and torch.is_grad_enabled() and (
                any(p.requires_grad for p in self.parameters()) or any(t.requires_grad for t in inputs)):
            if any(getattr(m, "with_norm", False) for m in self.modules()):
                raise NotImplementedError(
                    "training a neck with normalize=... means batch-statistics BatchNorm (the reference does not "
                    "freeze it): not on the B200 path; eval() runs with the norm folded into the conv epilogue")
            params = list(self.parameters())
            return tuple(training.PlanFunction.apply(self, len(inputs), *(list(inputs) + params)))
        return self._forward_infer(inputs)

    def _forward_infer(self, inputs, allow_split=True):
        want_fp32 = all(t.dtype == torch.float32 for t in inputs)
        # (the training path saves bf16 laterals for its backward plan: it never takes the split-precision path,
        # also when a frozen / eval backbone handed over fp32-I/O stage outputs)
        if allow_split and type(self) is FPN and not self._uses_gn() and \
                all(getattr(t, "_tdet_split", None) is not None for t in inputs):
            return self._forward_split([t._tdet_split for t in inputs])
        feats = [self._as_bf16_nhwc(t) for t in inputs]
        for t, c in zip(feats, self.in_channels):
            if t.dim() != 4 or t.shape[1] != c:
                raise ValueError("FPN input with %s channels, expected %d" % (tuple(t.shape), c))
        dev = feats[0].device
        operands = self._get_operands(dev)
        key = tuple(tuple(t.shape) for t in feats) + (dev,)
        entry = self._plans.get(key)
        if entry is None:
            entry = self._build_plan_gn(feats, operands) if self._uses_gn() else self._build_plan(feats, operands)
            self._plans[key] = entry
        plan, out_shapes = entry
        outs = [torch.empty(s, dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
                for s in out_shapes]
        plan.run(feats + outs)
        self._last_run = (plan, feats + outs)  # for profiling tools (bench.py)
        self._last_feats = feats
        self._last_key = key
        if want_fp32:
            outs = [o.float() for o in outs]
        return tuple(outs)

    def _forward_split(self, feats):
        """fp32-I/O mode: the backbone handed over split-precision stage outputs (bf16 hi|lo pairs)."""
        for t, c in zip(feats, self.in_channels):
            if t.dim() != 4 or t.shape[1] != 2 * c:
                raise ValueError("FPN split-precision input with %s channels, expected 2*%d" % (tuple(t.shape), c))
        dev = feats[0].device
        operands = self._get_operands(dev)
        key = ("split",) + tuple(tuple(t.shape) for t in feats) + (dev,)
        entry = self._plans.get(key)
        if entry is None:
            entry = self._build_plan(feats, operands, split=True)
            self._plans[key] = entry
        plan, split_shapes, f32_shapes = entry
        souts = [torch.empty(s, dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
                 for s in split_shapes]
        outs = [torch.empty(s, dtype=torch.float32, device=dev, memory_format=torch.channels_last)
                for s in f32_shapes]
        plan.run(list(feats) + souts + outs)
        self._last_run = (plan, list(feats) + souts + outs)
        for o, so in zip(outs, souts):
            o._tdet_split = so
        return tuple(outs)

    # ------------------------------------------------------------------ training path (config 4)
    def set_grad_sync(self, sync):
        """Attach a ``training.BucketAllReduce``: the neck's gradients form one flat fp32 bucket that is
        all-reduced as soon as the neck's backward is enqueued (it overlaps the whole backbone backward)."""
        self._grad_sync = sync
        if sync is not None:
            sync.attach(self, next(self.parameters()).device)

    def saved_activations(self):
        """{name: fp32 NCHW CPU tensor}: inputs ``C{j}`` and merged laterals ``lat{j}`` of the last training
        forward (test/debug API, see ResNet.saved_activations)."""
        state = self._train_state
        saved = {}
        used = state["feats"][self.start_level:self.backbone_end_level]
        for j, (t, l) in enumerate(zip(used, state["plan"].lats)):
            saved["C%d" % j] = t.detach().float().cpu()
            n, h, w, c = l.shape
            saved["lat%d" % j] = l.buf[:n * h * w * c].view(n, h, w, c).permute(0, 3, 1, 2).float().cpu()
        return saved

    def _train_forward(self, inputs, params):
        with torch.no_grad():
            outs = self._forward_infer(inputs, allow_split=False)
        plan, ext = self._last_run
        self._train_serial = getattr(self, "_train_serial", 0) + 1
        state = dict(plan=plan, feats=self._last_feats, outs=list(ext[len(self._last_feats):]),
                     in_dtypes=[t.dtype for t in inputs], params=list(params),
                     serial=self._train_serial, key=self._last_key)
        self._train_state = state
        return outs, state

    def _build_bwd_plan(self, state, operands):
        plan, feats, outs = state["plan"], state["feats"], state["outs"]
        dev = feats[0].device
        used = feats[self.start_level:self.backbone_end_level]
        nl = len(used)
        bb = training.BackwardBuilder(dev, operands)
        convs = [cm.conv for _, group in self._conv_groups() for cm in group]
        plist = []
        for c in convs:
            plist.append(c.weight)
            if c.bias is not None:
                plist.append(c.bias)
        bucket = training.GradBucket(plist, dev)
        g_ext = [engine.nhwc_empty(o.shape[0], o.shape[2], o.shape[3], o.shape[1], dev) for o in outs]
        gp = [engine.act_of(g) for g in g_ext]
        extra_dc = None  # gradient of the extra stride-2 convs w.r.t. the top backbone level
        out_acts = [engine.act_of(o) for o in outs]  # the forward's returned tensors (external, re-bound)
        if not self.add_extra_convs:
            # extra levels are stride-2 subsamples of the last output (fpn.py:114-116): scatter them back
            for j in range(len(outs) - 1, nl - 1, -1):
                up = bb.new_act(gp[j - 1].shape)
                bb.ops.append(engine.op_dilate2(gp[j], up))
                tot = bb.new_act(gp[j - 1].shape)
                bb.ops.append(engine.op_add_mask(gp[j - 1], tot, residual=up))
                bb.release(up)
                gp[j - 1] = tot
        elif len(outs) > nl:
            # RetinaNet-style extra levels (fpn.py:118-124): o_nl = conv(C_top), o_j = conv(relu_(o_{j-1})); every
            # level that feeds another one is RETURNED post-ReLU (in place), so its incoming gradient and the
            # next conv's data gradient are merged and masked in that dgrad conv's epilogue
            top = engine.act_of(feats[self.backbone_end_level - 1])
            g = gp[len(outs) - 1]
            for j in range(len(outs) - 1, nl - 1, -1):
                conv = self.fpn_convs[j].conv
                x_j = top if j == nl else out_acts[j - 1]
                bb.wgrad("out%d" % j, conv, None, x_j, g, bucket.view(bucket.index_of(conv.weight)))
                if conv.bias is not None:
                    bb.ops.append(engine.op_colsum(g, bucket.view(bucket.index_of(conv.bias))))
                if j > nl:
                    g_prev = bb.dgrad("out%d" % j, conv, None, g, x_j.shape, residual=gp[j - 1], mask=out_acts[j - 1])
                else:
                    g_prev = bb.dgrad("out%d" % j, conv, None, g, x_j.shape)
                if g is not gp[len(outs) - 1]:
                    bb.release(g)
                g = g_prev
            extra_dc = g
        # (subclasses whose returned levels are not the output convs' own results map the gradients back here)
        gp = self._bwd_to_pyramid(bb, bucket, state, gp, out_acts, nl)
        d_feats = [engine.nhwc_empty(t.shape[0], t.shape[2], t.shape[3], t.shape[1], dev) for t in used]
        pooled = None
        for j in range(nl):
            lat, out = self.lateral_convs[j].conv, self.fpn_convs[j].conv
            L = plan.lats[j]
            # output conv: dW, db, and dL_j = dgrad(dP_j) + 2x2 sum-pool of the finer level's dL
            bb.wgrad("out%d" % j, out, None, L, gp[j], bucket.view(bucket.index_of(out.weight)))
            if out.bias is not None:
                bb.ops.append(engine.op_colsum(gp[j], bucket.view(bucket.index_of(out.bias))))
            dL = bb.dgrad("out%d" % j, out, None, gp[j], L.shape, residual=pooled)
            bb.release(pooled)
            # lateral conv
            bb.wgrad("lat%d" % j, lat, None, engine.act_of(used[j]), dL,
                     bucket.view(bucket.index_of(lat.weight)))
            if lat.bias is not None:
                bb.ops.append(engine.op_colsum(dL, bucket.view(bucket.index_of(lat.bias))))
            wd = bb.dgrad_weight("lat%d" % j, lat, None)
            bb.ops.append(engine.op_conv(dL, wd, engine.act_of(d_feats[j]), 1, 1, 1, 0, 1,
                                         residual=extra_dc if (j == nl - 1 and extra_dc is not None) else None))
            if j + 1 < nl:
                pooled = bb.new_act(plan.lats[j + 1].shape)
                bb.ops.append(engine.op_sumpool2(dL, pooled))
            else:
                pooled = None
            bb.release(dL)
        ops, _ = bb.finalize()
        ops = [engine.op_zero(bucket.flat)] + ops
        ext = g_ext + list(feats) + d_feats + list(outs)
        bplan = engine.Plan(ops, ext, [operands, bb.buffers, bb.acc_ws, bucket.flat], dev)
        return bplan, bucket

    def _bwd_to_pyramid(self, bb, bucket, state, gp, out_acts, nl):
        """Gradients w.r.t. the output convs' results P_0..P_{nl-1}, given those w.r.t. the returned levels: the
        same thing for an FPN."""
        return gp

    def _train_backward(self, state, gouts):
        if state["serial"] != getattr(self, "_train_serial", 0):
            raise RuntimeError("FPN backward after a newer training forward of the same module "
                               "(one forward per backward: saved laterals live in the plan)")
        feats, outs = state["feats"], state["outs"]
        dev = feats[0].device
        operands = self._get_operands(dev)
        key = ("bwd", id(state["plan"]))
        entry = self._plans.get(key, group=state["key"])
        if entry is None:
            entry = self._plans.put(key, self._build_bwd_plan(state, operands), group=state["key"])
        bplan, bucket = entry
        gs = [training.as_grad_nhwc(g, o) for g, o in zip(gouts, outs)]
        used = feats[self.start_level:self.backbone_end_level]
        d_feats = [torch.empty_like(t, memory_format=torch.channels_last) for t in used]
        ext = gs + list(feats) + d_feats + list(outs)
        sync = getattr(self, "_grad_sync", None)
        if sync is not None:
            sync.guard(bucket.flat)
        bplan.run(ext)
        self._last_bwd_run = (bplan, ext)
        if sync is not None:
            flat = sync.reduce(bucket.flat, bucket.params)
            sync.module_done()
        else:
            flat = bucket.flat.clone()
        lookup = {id(p): flat[bucket.offsets[i]:bucket.offsets[i] + p.numel()].view(p.shape)
                  for i, p in enumerate(bucket.params)}
        g_params = [lookup[id(p)] for p in state["params"]]
        g_inputs = [None] * len(feats)
        for j, d in enumerate(d_feats):
            i = self.start_level + j
            g_inputs[i] = d if state["in_dtypes"][i] == torch.bfloat16 else d.to(state["in_dtypes"][i])
        return g_inputs, g_params


def _epi(operands, key):
    """Epilogue operands of conv `key`: folded BatchNorm scale/shift if it has a norm layer, else its bias."""
    bn = operands.value(key + ".bn")
    if bn is not None:
        return dict(scale=bn[0], shift=bn[1])
    return dict(shift=operands.value(key + ".b"))


def _fold_conv_bn(conv, norm, out):
    """(scale, shift) of conv(+bias) -> eval BatchNorm: y = scale * conv_nobias(x) + shift."""
    sc, sh = engine.fold_bn(norm, out=out)
    if conv.bias is not None:
        sh.add_(conv.bias.detach().float() * sc)   # BN(conv + b) = scale*conv + (shift + scale*b)
    return sc, sh


def _gn_affine(norm, out):
    g, b = norm.weight.detach().float(), norm.bias.detach().float()
    if out is None:
        return g.contiguous().clone(), b.contiguous().clone()
    out[0].copy_(g)
    out[1].copy_(b)
    return out


def _bias_copy(bias, out):
    src = bias.detach().float()
    if out is None:
        return src.contiguous().clone()
    out.copy_(src)
    return out
