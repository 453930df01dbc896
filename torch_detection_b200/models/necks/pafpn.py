"""Path-Aggregation FPN neck, B200 execution -- drop-in for reference ``models/necks/pafpn.py``
(SURVEY.md 8(f) row f3).

Constructor arguments, attributes (``lateral_convs`` / ``fpn_convs`` / ``pa_convs1`` / ``pa_convs2``),
``state_dict`` keys, ``init_weights`` and the returned tuple are those of the reference ``PAFPN``
(pafpn.py:9-148).  On top of the FPN plan (laterals with the fused top-down add, 3x3 smoothing convs)
the bottom-up path of pafpn.py:131-134

    N_0 = P_0 ;  N_i = pa_convs2[i-1]( P_i + pa_convs1[i-1](N_{i-1}) )

runs as two fused convs per level: the stride-2 3x3 ``pa_convs1`` takes P_i as its residual operand
(bias + add in the fp32 epilogue, so the sum is written once), then the 3x3 ``pa_convs2``.  With
``activation='relu'`` the ReLU sits between conv and add (pafpn.py:63-80 -> ConvModule), so the add is a
separate fused add op.  Training runs as one ``autograd.Function`` like the FPN's (``_bwd_to_pyramid`` maps the
gradients of the returned levels back to the pyramid); ``activation='relu6'`` is inference only.
"""
import torch
import torch.nn as nn

from ... import engine
from ...registry import NECKS
from ..utils import ConvModule
from .fpn import FPN, _epi


@NECKS.register_module
class PAFPN(FPN):

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1,
                 add_extra_convs=False, normalize=None, use_gn=False, activation=None):
        if activation not in (None, "relu", "relu6"):
            raise ValueError("PAFPN activation must be None, 'relu' or 'relu6'")
        nn.Module.__init__(self)
        assert isinstance(in_channels, list)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_ins = len(in_channels)
        self.num_outs = num_outs
        self.with_bias = normalize is None
        self.activation = activation
        if end_level == -1:
            self.backbone_end_level = self.num_ins
            assert num_outs >= self.backbone_end_level - start_level
        else:
            self.backbone_end_level = end_level
            assert end_level <= self.num_ins
            assert num_outs == end_level - start_level
        self.start_level = start_level
        self.end_level = end_level
        self.add_extra_convs = add_extra_convs

        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        self.pa_convs1 = nn.ModuleList()
        self.pa_convs2 = nn.ModuleList()
        # creation order = the reference's (pafpn.py:46-80): parameter construction consumes the RNG alike
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(ConvModule(in_channels[i], out_channels, 1, normalize=normalize,
                                                 bias=self.with_bias, use_gn=use_gn))
            self.fpn_convs.append(ConvModule(out_channels, out_channels, 3, padding=1, normalize=normalize,
                                             bias=self.with_bias, use_gn=use_gn))
            if i < self.backbone_end_level - 1:
                self.pa_convs1.append(_pa_conv(out_channels, 2, normalize, self.with_bias, use_gn, activation))
                self.pa_convs2.append(_pa_conv(out_channels, 1, normalize, self.with_bias, use_gn, activation))
        extra_levels = num_outs - self.backbone_end_level + self.start_level
        if add_extra_convs and extra_levels >= 1:
            for i in range(extra_levels):
                cin = in_channels[self.backbone_end_level - 1] if i == 0 else out_channels
                self.fpn_convs.append(ConvModule(cin, out_channels, 3, stride=2, padding=1, normalize=normalize,
                                                 bias=self.with_bias, use_gn=use_gn))
        self._plans = engine.PlanCache()
        self._operands = None
        self._operand_key = None

    def _conv_groups(self):
        return (("lat", self.lateral_convs), ("out", self.fpn_convs), ("pa1_", self.pa_convs1),
                ("pa2_", self.pa_convs2))

    def _emit_outputs(self, ops, operands, lats, shapes, dev):
        co = self.out_channels
        relu = self.activation is not None
        relu6 = self.activation == "relu6"
        nl = len(lats)
        keep = []

        def temp(shape):
            t = torch.empty(shape[0] * shape[1] * shape[2] * shape[3], dtype=torch.bfloat16, device=dev)
            keep.append(t)
            return engine.Act(t, shape, torch.bfloat16)

        outs = []
        prev = None
        saved = dict(P=[None] * nl, s=[None] * nl, t=[None] * nl)   # what the backward plan reads (training)
        self._pa_saved = saved
        for j in range(nl):
            nb, h, w, _ = shapes[j]
            # P_j: the returned N_0 for the finest level, a temporary otherwise
            p = engine.act_of(engine.nhwc_empty(nb, h, w, co, dev)) if j == 0 else temp((nb, h, w, co))
            ops.append(engine.op_conv(lats[j], operands.value("out%d.w" % j), p, 3, 3, 1, 1, 1,
                                      **_epi(operands, "out%d" % j)))
            saved["P"][j] = p
            if j == 0:
                prev = p
                outs.append(p)
                continue
            if prev.shape[1] != 2 * h and engine.conv_out(prev.shape[1], 3, 2, 1) != h:
                raise RuntimeError("PAFPN level sizes do not match")
            s = temp((nb, h, w, co))
            if relu:
                # relu(conv + b) first, then + P_j (ConvModule applies the activation before the add)
                t = temp((nb, h, w, co))
                ops.append(engine.op_conv(prev, operands.value("pa1_%d.w" % (j - 1)), t, 3, 3, 2, 1, 1,
                                          **_epi(operands, "pa1_%d" % (j - 1)), relu=True, relu6=relu6))
                ops.append(engine.op_add_mask(t, s, residual=p))
                saved["t"][j] = t
            else:
                ops.append(engine.op_conv(prev, operands.value("pa1_%d.w" % (j - 1)), s, 3, 3, 2, 1, 1,
                                          **_epi(operands, "pa1_%d" % (j - 1)), residual=p))
            saved["s"][j] = s
            nj = engine.act_of(engine.nhwc_empty(nb, h, w, co, dev))
            ops.append(engine.op_conv(s, operands.value("pa2_%d.w" % (j - 1)), nj, 3, 3, 1, 1, 1,
                                      **_epi(operands, "pa2_%d" % (j - 1)), relu=relu, relu6=relu6))
            outs.append(nj)
            prev = nj
        return outs, keep

    def _build_plan(self, feats, operands, split=False):
        entry = FPN._build_plan(self, feats, operands, split=split)
        entry[0].pa = self._pa_saved   # P_j, s_j = P_j + act(pa1(N_{j-1})), t_j = act(pa1(N_{j-1})): saved activations
        return entry

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        if self.training and torch.is_grad_enabled() and (
                any(p.requires_grad for p in self.parameters()) or any(t.requires_grad for t in inputs)):
            if self.activation == "relu6":
                raise NotImplementedError("PAFPN(activation='relu6') training is not on the B200 path: the backward "
                                          "kernels' mask operand encodes ReLU only (inference runs)")
        return FPN.forward(self, inputs)

    def _bwd_to_pyramid(self, bb, bucket, state, gp, out_acts, nl):
        """Backward of the bottom-up path (pafpn.py:131-134), top level first:
            N_j = act(pa2(s_j)),  s_j = P_j + t_j,  t_j = act(pa1(N_{j-1}))        (N_0 = P_0)
        gp[j] arrives as dL/dN_j from outside (extra levels already folded in) and leaves as dL/dP_j.  The ReLU
        backward of N_j rides in the epilogue of the stride-2 dgrad that delivers level j + 1's contribution (mask
        operand), the one of t_j is one masked copy."""
        saved = state["plan"].pa
        relu = self.activation == "relu"
        gP = list(gp)
        carry = None   # dL/dN_j complete (outside + level j + 1) and already masked with N_j > 0
        for j in range(nl - 1, 0, -1):
            pa1, pa2 = self.pa_convs1[j - 1].conv, self.pa_convs2[j - 1].conv
            if carry is not None:
                g2 = carry
            elif relu:
                g2 = bb.new_act(gp[j].shape)
                bb.ops.append(engine.op_add_mask(gp[j], g2, mask=out_acts[j]))
            else:
                g2 = gp[j]
            s_j, t_j = saved["s"][j], saved["t"][j]
            bb.wgrad("pa2_%d" % (j - 1), pa2, None, s_j, g2, bucket.view(bucket.index_of(pa2.weight)))
            if pa2.bias is not None:
                bb.ops.append(engine.op_colsum(g2, bucket.view(bucket.index_of(pa2.bias))))
            g_s = bb.dgrad("pa2_%d" % (j - 1), pa2, None, g2, s_j.shape)
            if g2 is not gp[j]:
                bb.release(g2)
            gP[j] = g_s                      # s_j = P_j + t_j
            if relu:
                g_t = bb.new_act(g_s.shape)
                bb.ops.append(engine.op_add_mask(g_s, g_t, mask=t_j))
            else:
                g_t = g_s
            n_prev = out_acts[j - 1]
            bb.wgrad("pa1_%d" % (j - 1), pa1, None, n_prev, g_t, bucket.view(bucket.index_of(pa1.weight)))
            if pa1.bias is not None:
                bb.ops.append(engine.op_colsum(g_t, bucket.view(bucket.index_of(pa1.bias))))
            carry = bb.dgrad("pa1_%d" % (j - 1), pa1, None, g_t, n_prev.shape, residual=gp[j - 1],
                             mask=out_acts[j - 1] if (relu and j - 1 > 0) else None)
            if g_t is not g_s:
                bb.release(g_t)
        if carry is not None:
            gP[0] = carry
        return gP


def _pa_conv(channels, stride, normalize, bias, use_gn, activation):
    """3x3 ConvModule of the bottom-up path; the activation is applied by the owning plan."""
    return ConvModule(channels, channels, 3, stride=stride, padding=1, normalize=normalize, bias=bias,
                      use_gn=use_gn, activation=activation)
