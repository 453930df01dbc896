"""Host-side glue between torch tensors (device memory, streams) and the C ABI.

torch is plumbing here: it owns device memory and the current stream.  Every arithmetic step of the
path runs in ``libtdet_b200.so``; nothing in this module computes on tensors with torch ops.
"""
import collections
import ctypes
import os

import torch

from . import _C

BN_EPS = 1e-5  # nn.BatchNorm2d default (reference models/utils/layers.py:50-54)

_TD = {torch.bfloat16: _C.BF16, torch.float16: _C.F16, torch.float32: _C.F32, torch.uint8: _C.U8}


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _index(device):
    return device.index if device.index is not None else torch.cuda.current_device()


def require_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise NotImplementedError(
            "%s must be a CUDA tensor: the B200 path has no CPU fallback" % what)


def conv_out(v, k, s, p, d=1):
    return (v + 2 * p - d * (k - 1) - 1) // s + 1


def set_sm_reserve(device, sms):
    """Plans built afterwards leave `sms` SMs free (for NCCL kernels overlapping the backward pass)."""
    _C.check(_C.lib().tdet_set_sm_reserve(_index(device), sms))


def stem_staging_dims(ho, wo):
    """(hp, wp) of the padded NHWC4 staging buffer TDET_OP_PREP writes for a stem output ho x wo."""
    hp, wp = ctypes.c_int32(), ctypes.c_int32()
    _C.check(_C.lib().tdet_stem_staging_dims(ho, wo, ctypes.byref(hp), ctypes.byref(wp)))
    return hp.value, wp.value


class Act(object):
    """Handle of a dense NHWC 16-bit activation tensor: device buffer, logical (n, h, w, c) shape,
    storage dtype and the device address of its ``tdet_tensor_meta`` (None = plain values)."""
    __slots__ = ("buf", "shape", "dtype", "meta", "offset")

    def __init__(self, buf, shape, dtype=None, meta=None, offset=0):
        self.buf = buf
        self.shape = tuple(shape)
        self.dtype = dtype if dtype is not None else buf.dtype
        self.meta = meta
        self.offset = offset  # in elements, from the start of `buf`

    @property
    def ptr(self):
        return self.buf.data_ptr() + 2 * self.offset


def act_of(t, meta=None):
    """Act view of a logical-NCHW channels_last tensor (zero copy)."""
    n, c, h, w = t.shape
    return Act(t, (n, h, w, c), t.dtype, meta)


class MetaArena(object):
    """Contiguous array of ``tdet_tensor_meta`` (8 bytes each) in device memory."""

    def __init__(self, count, device):
        self.tensor = torch.zeros((max(count, 1), 2), dtype=torch.int32, device=device)
        self.count = count
        self._next = 0

    def new(self):
        assert self._next < self.count
        addr = self.tensor.data_ptr() + 8 * self._next
        self._next += 1
        return addr

    def read(self):
        """[(exponent, amax)] -- synchronises; debugging / tests only."""
        raw = self.tensor.cpu()
        amax = raw[:, 1].contiguous().view(torch.float32)
        return [(int(raw[i, 0]), float(amax[i])) for i in range(self._next)]


# --------------------------------------------------------------------------------------------
# operand preparation
# --------------------------------------------------------------------------------------------

class OperandCache(object):
    """Derived device operands (packed weights, folded BN, bound constants) of one module.

    Every entry remembers how it is made and which parameters / buffers it depends on; `refresh()`
    re-derives, IN PLACE, the entries whose dependencies changed (an optimizer step, a
    load_state_dict), so device pointers -- and with them the compiled plans and their TMA
    descriptors -- stay valid for the life of the module."""

    def __init__(self):
        self.store = {}

    @staticmethod
    def _versions(deps):
        return tuple((p.data_ptr(), p._version) for p in deps)

    def get(self, key, make, deps=()):
        """Existing entry, or make(None) (remembered together with how to re-derive it in place)."""
        e = self.store.get(key)
        if e is None:
            e = [make(None), make, tuple(deps), self._versions(deps)]
            self.store[key] = e
        return e[0]

    def value(self, key):
        """Existing entry or None (optional operands such as a conv bias)."""
        e = self.store.get(key)
        return None if e is None else e[0]

    def refresh(self, force=False):
        """Re-derives the entries whose dependencies changed.  Changes are detected through
        ``(data_ptr, tensor._version)``; in-place updates made THROUGH ``.data`` (``p.data.add_``: legacy
        optimizers, EMA, weight clamping) do not bump ``_version`` -- after such an update call
        ``module.invalidate_operands()`` (or set ``TDET_REFRESH_EVERY_STEP=1``), which passes ``force=True``
        and re-derives every entry that depends on a parameter or buffer."""
        n = 0
        for e in self.store.values():
            v = self._versions(e[2])
            if v != e[3] or (force and e[2]):
                e[1](e[0])
                e[3] = v
                n += 1
        return n


REFRESH_EVERY_STEP = os.environ.get("TDET_REFRESH_EVERY_STEP", "0") != "0"


class PlanCache(object):
    """Bounded LRU of compiled plans.  Entries are grouped (a training forward plan and its backward plan share a
    group and are evicted together); at most ``capacity`` groups stay alive (``TDET_PLAN_CACHE``, default 4).  A plan
    pins its whole static activation arena -- for training every saved activation plus the gradient pool -- and
    the reference's loader pads per batch, so detection training can see many distinct H x W: an unbounded cache
    would grow by gigabytes per new shape until the allocator fails.  Evicted plans are rebuilt on demand (fixed
    padded shapes avoid the rebuilds)."""

    def __init__(self, capacity=None):
        self.capacity = capacity if capacity is not None else max(1, int(os.environ.get("TDET_PLAN_CACHE", "4")))
        self.groups = collections.OrderedDict()

    def get(self, key, group=None):
        g = key if group is None else group
        d = self.groups.get(g)
        if d is None or key not in d:
            return None
        self.groups.move_to_end(g)
        return d[key]

    def put(self, key, entry, group=None):
        g = key if group is None else group
        self.groups.setdefault(g, {})[key] = entry
        self.groups.move_to_end(g)
        while len(self.groups) > self.capacity:
            self.groups.popitem(last=False)
        return entry

    def __setitem__(self, key, entry):
        self.put(key, entry)

    def __len__(self):
        return sum(len(d) for d in self.groups.values())

    def clear(self):
        self.groups.clear()


def pack_conv_weight(w, dtype=torch.bfloat16, out=None):
    """fp32 OIHW parameter -> 16-bit [O][kh][kw][I] (tcgen05 B operand rows) in `dtype`."""
    require_cuda(w, "weight")
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    o, i, kh, kw = w.shape
    if out is None:
        out = torch.empty((o, kh, kw, i), dtype=dtype, device=w.device)
    with torch.cuda.device(w.device):
        _C.check(_C.lib().tdet_pack_conv_weight(w.data_ptr(), out.data_ptr(), o, i, kh, kw,
                                                _TD[dtype], _stream_ptr(w.device)))
    return out


def dense_group_weight(w, groups):
    """[cout][cin / groups][kh][kw] parameter of a grouped conv -> its dense block-diagonal [cout][cin][kh][kw] form
    (operand preparation when the weights change; the training path derives its dgrad operand from it)."""
    w = w.detach().float()
    cout, cig, kh, kw = w.shape
    cog = cout // groups
    dense = w.new_zeros(groups, cog, groups, cig, kh, kw)
    idx = torch.arange(groups, device=w.device)
    dense[idx, :, idx] = w.view(groups, cog, cig, kh, kw)
    return dense.view(cout, groups * cig, kh, kw)


def pack_dual_weight(w, scale, w2, scale2, dtype=torch.bfloat16, out=None):
    """[cout][cin + cin2] 16-bit operand of a dual-source 1x1 conv: [scale * w | scale2 * w2] (each BatchNorm scale
    folded into its half before the one rounding)."""
    require_cuda(w, "weight")
    ws = [t.detach().float().contiguous() for t in (w, w2)]
    cout, cin = ws[0].shape[0], ws[0].shape[1]
    cin2 = ws[1].shape[1]
    assert ws[0].shape[2:] == (1, 1) and ws[1].shape[2:] == (1, 1) and ws[1].shape[0] == cout
    if out is None:
        out = torch.empty((cout, cin + cin2), dtype=dtype, device=w.device)
    es = out.element_size()
    with torch.cuda.device(w.device):
        _C.check(_C.lib().tdet_pack_conv_weight_scaled(ws[0].data_ptr(), _ptr(scale), out.data_ptr(), cout, cin, 1, 1,
                                                       cin + cin2, _TD[dtype], _stream_ptr(w.device)))
        _C.check(_C.lib().tdet_pack_conv_weight_scaled(ws[1].data_ptr(), _ptr(scale2), out.data_ptr() + es * cin, cout,
                                                       cin2, 1, 1, cin + cin2, _TD[dtype], _stream_ptr(w.device)))
    return out


def pack_conv_weight_split(w, out=None):
    """fp32 OIHW -> bf16 [O][kh][kw][2*I]: per tap the hi halves bf16(w), then the lo halves bf16(w - hi)."""
    require_cuda(w, "weight")
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    o, i, kh, kw = w.shape
    if out is None:
        out = torch.empty((o, kh, kw, 2 * i), dtype=torch.bfloat16, device=w.device)
    with torch.cuda.device(w.device):
        _C.check(_C.lib().tdet_pack_conv_weight_split(w.data_ptr(), out.data_ptr(), o, i, kh, kw,
                                                      _stream_ptr(w.device)))
    return out


def pack_stem_weight_split(w, out=None):
    """fp32 [64][3][7][7] -> bf16 [64][896]: the stem operand layout for hi, then for lo."""
    require_cuda(w, "weight")
    w = w.detach().float().contiguous()
    if tuple(w.shape) != (64, 3, 7, 7):
        raise ValueError("stem weight must be (64,3,7,7)")
    if out is None:
        out = torch.empty((64, 896), dtype=torch.bfloat16, device=w.device)
    with torch.cuda.device(w.device):
        _C.check(_C.lib().tdet_pack_stem_weight_split(w.data_ptr(), out.data_ptr(), _stream_ptr(w.device)))
    return out


def pack_grouped_conv_weight(w, groups, dtype=torch.bfloat16, out=None):
    """fp32 [O][I/groups][kh][kw] grouped parameter -> dense block-diagonal 16-bit [O][kh][kw][I]."""
    require_cuda(w, "weight")
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    o, ig, kh, kw = w.shape
    i = ig * groups
    if out is None:
        out = torch.empty((o, kh, kw, i), dtype=dtype, device=w.device)
    with torch.cuda.device(w.device):
        _C.check(_C.lib().tdet_pack_grouped_conv_weight(w.data_ptr(), out.data_ptr(), o, i, kh, kw, groups,
                                                        _TD[dtype], _stream_ptr(w.device)))
    return out


def pack_dgrad_weight(w, scale=None, dtype=torch.bfloat16, out=None):
    """fp32 OIHW parameter (+ folded BN scale per O) -> 16-bit [I][kh][kw][O] with the filter rotated by
    180 degrees: the B operand of the data-gradient conv (a conv over the output gradient)."""
    require_cuda(w, "weight")
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    o, i, kh, kw = w.shape
    if out is None:
        out = torch.empty((i, kh, kw, o), dtype=dtype, device=w.device)
    with torch.cuda.device(w.device):
        _C.check(_C.lib().tdet_pack_dgrad_weight(w.data_ptr(), _ptr(scale), out.data_ptr(), o, i, kh, kw,
                                                 _TD[dtype], _stream_ptr(w.device)))
    return out


def pack_stem_weight(w, out=None, dtype=torch.bfloat16):
    """fp32 [64][3][7][7] -> 16-bit [64][448] stem operand (in the staging's format)."""
    require_cuda(w, "weight")
    w = w.detach()
    if tuple(w.shape) != (64, 3, 7, 7):
        raise ValueError("stem weight must be (64,3,7,7)")
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    if out is None:
        out = torch.empty((64, 448), dtype=dtype, device=w.device)
    with torch.cuda.device(w.device):
        _C.check(_C.lib().tdet_pack_stem_weight(w.data_ptr(), out.data_ptr(), _TD[out.dtype], _stream_ptr(w.device)))
    return out


def fold_bn(bn, out=None):
    """eval-mode nn.BatchNorm2d -> (scale, shift) fp32 device vectors."""
    g, b, m, v = (t.detach().float().contiguous() for t in
                  (bn.weight, bn.bias, bn.running_mean, bn.running_var))
    require_cuda(g, "BatchNorm parameters")
    ch = g.numel()
    if out is None:
        scale = torch.empty(ch, dtype=torch.float32, device=g.device)
        shift = torch.empty(ch, dtype=torch.float32, device=g.device)
    else:
        scale, shift = out
    with torch.cuda.device(g.device):
        _C.check(_C.lib().tdet_fold_bn(g.data_ptr(), b.data_ptr(), m.data_ptr(), v.data_ptr(),
                                       ctypes.c_float(bn.eps), scale.data_ptr(), shift.data_ptr(),
                                       ch, _stream_ptr(g.device)))
    return scale, shift


def bound_consts(w_packed, scale, shift, out=None):
    """{G, max|shift|} with |conv(x)*scale + shift| <= G*max|x| + max|shift| (device, fp32[2])."""
    cout = w_packed.shape[0]
    k = w_packed.numel() // cout
    if out is None:
        out = torch.empty(2, dtype=torch.float32, device=w_packed.device)
    with torch.cuda.device(w_packed.device):
        _C.check(_C.lib().tdet_conv_bound_consts(w_packed.data_ptr(), _TD[w_packed.dtype], _ptr(scale),
                                                 _ptr(shift), cout, k, out.data_ptr(),
                                                 _stream_ptr(w_packed.device)))
    return out


# --------------------------------------------------------------------------------------------
# op descriptors
# --------------------------------------------------------------------------------------------

def _ptr(t):
    return None if t is None else t.data_ptr()


def nhwc_empty(n, h, w, c, device, dtype=torch.bfloat16):
    """Dense NHWC 16-bit buffer exposed as a logical-NCHW channels_last tensor (zero-copy view)."""
    return torch.empty((n, c, h, w), dtype=dtype, device=device, memory_format=torch.channels_last)


def op_conv(x, wgt, y, kh, kw, stride, pad, dil=1, scale=None, shift=None, residual=None,
            coarse=None, relu=False, consts=None, scaled_out=False, mask=None, coarse_parity=False, groups=1,
            split=False, dual=None, relu6=False, reverse=False):
    """x, y, residual, coarse, mask: ``Act`` handles; wgt packed [cout][kh][kw][cin] in x's dtype.
    reverse: the kernel walks its tiles from the last to the first (TDET_FLAG_REVERSE; same results -- plans set it
    on every second conv by themselves).
    split: split-precision tensors (bf16 hi|lo pairs, 2x the logical channels in memory); Act shapes stay
    logical.
    dual: (x2 Act, stride2) -- second input of a 1x1 conv (TDET_FLAG_DUAL: the projection shortcut of a stage's
    first bottleneck contracted in the same launch); wgt is then the K-concatenation [cout][cin + cin2] from
    ``pack_dual_weight`` and x, x2 are plain tensors of one format."""
    n, h, w, cin = x.shape
    cout = wgt.shape[0]
    if dual is not None and wgt.numel() != cout * (cin + dual[0].shape[3]):
        raise ValueError("dual-source conv: weights must be the [cout][cin + cin2] concatenation")
    if wgt.dtype != x.dtype:
        raise ValueError("conv weights must be packed in the input tensor's format (%s vs %s)"
                         % (wgt.dtype, x.dtype))
    op = _C.TdetOp()
    op.kind = _C.OP_CONV
    op.flags = (_C.FLAG_RELU6 if relu6 else (_C.FLAG_RELU if relu else 0)) | \
        (_C.FLAG_SCALED_OUT if scaled_out else 0) | \
        (_C.FLAG_COARSE_PARITY if coarse_parity else 0) | (_C.FLAG_SPLIT if split else 0) | \
        (_C.FLAG_REVERSE if reverse else 0)
    if mask is not None:
        op.mask = mask.ptr
    op.groups = groups
    op.n, op.h, op.w, op.cin = n, h, w, cin
    op.cout, op.kh, op.kw = cout, kh, kw
    op.stride, op.pad, op.dil = stride, pad, dil
    op.ho, op.wo = conv_out(h, kh, stride, pad, dil), conv_out(w, kw, stride, pad, dil)
    assert y.shape == (n, op.ho, op.wo, cout), (y.shape, (n, op.ho, op.wo, cout))
    op.x_dtype, op.y_dtype = _TD[x.dtype], _TD[y.dtype]
    op.x, op.wgt, op.y = x.ptr, wgt.data_ptr(), y.ptr
    op.x_meta, op.y_meta = x.meta, y.meta
    op.scale, op.shift = _ptr(scale), _ptr(shift)
    if residual is not None:
        op.residual, op.residual_dtype, op.residual_meta = residual.ptr, _TD[residual.dtype], residual.meta
    if coarse is not None:
        op.coarse, op.coarse_dtype, op.coarse_meta = coarse.ptr, _TD[coarse.dtype], coarse.meta
        op.hc, op.wc = coarse.shape[1], coarse.shape[2]
    op.bound_consts = _ptr(consts)
    if dual is not None:
        x2, stride2 = dual
        if x2.dtype != x.dtype:
            raise ValueError("dual-source conv: both inputs share one 16-bit format")
        op.flags |= _C.FLAG_DUAL
        op.x2, op.x2_meta, op.x2_dtype = x2.ptr, x2.meta, _TD[x2.dtype]
        op.cin2, op.stride2, op.h2, op.w2 = x2.shape[3], stride2, x2.shape[1], x2.shape[2]
    return op


def op_bottleneck_tail(x, w2, y, residual, w3, bn2, bn3, consts2=None, consts3=None, scaled_out=False, nxt=None):
    """TDET_OP_BOTTLENECK_TAIL: z2 = relu(bn2(conv3x3(x))), y = relu(bn3(conv1x1(z2)) + residual) and -- with
    nxt = dict(w=[64][256] in y's dtype, bn=(scale, shift), y=Act, consts=, scaled_out=) -- the next block's
    y2 = relu(bn1(conv1x1(y))), in one kernel.  x: Act [n][h][w][64]; w2 [64][3][3][64], w3 [256][64] in x's dtype;
    bn2 / bn3: (scale, shift) fp32 device vectors."""
    n, h, w, cin = x.shape
    if w2.dtype != x.dtype or w3.dtype != x.dtype:
        raise ValueError("bottleneck tail: conv2 / conv3 weights must be packed in the input tensor's format")
    op = _C.TdetOp()
    op.kind = _C.OP_BOTTLENECK_TAIL
    op.flags = _C.FLAG_RELU | (_C.FLAG_SCALED_OUT if scaled_out else 0)
    op.n, op.h, op.w, op.cin = n, h, w, cin
    op.cout, op.cout2, op.kh, op.kw = w3.shape[0], w2.shape[0], 3, 3
    op.stride, op.pad, op.dil = 1, 1, 1
    op.ho, op.wo = h, w
    assert y.shape == (n, h, w, op.cout) and residual.shape == y.shape
    op.x_dtype, op.y_dtype, op.residual_dtype = _TD[x.dtype], _TD[y.dtype], _TD[residual.dtype]
    op.x, op.wgt, op.y, op.residual = x.ptr, w2.data_ptr(), y.ptr, residual.ptr
    op.x_meta, op.y_meta, op.residual_meta = x.meta, y.meta, residual.meta
    op.scale, op.shift = _ptr(bn2[0]), _ptr(bn2[1])
    op.wgt2, op.scale2, op.shift2 = w3.data_ptr(), _ptr(bn3[0]), _ptr(bn3[1])
    op.bound_consts, op.bound_consts2 = _ptr(consts2), _ptr(consts3)
    if nxt is not None:
        y2 = nxt["y"]
        if nxt["w"].dtype != y.dtype:
            raise ValueError("bottleneck tail: the next conv1's weights must be packed in y's format")
        assert y2.shape == (n, h, w, nxt["w"].shape[0])
        op.wgt3, op.scale3, op.shift3 = nxt["w"].data_ptr(), _ptr(nxt["bn"][0]), _ptr(nxt["bn"][1])
        op.bound_consts3 = _ptr(nxt.get("consts"))
        op.y2, op.y2_meta, op.y2_dtype, op.cout3 = y2.ptr, y2.meta, _TD[y2.dtype], nxt["w"].shape[0]
        if nxt.get("scaled_out"):
            op.flags |= _C.FLAG_SCALED_OUT2
    return op


def op_prep(x, y, ho, wo, y_meta=None, scale=None, shift=None, padded_hw=None, split=False, y_dtype=torch.bfloat16):
    """x: logical (n, 3, h, w) image batch (fp32 / bf16 / uint8, any strides: an HWC batch viewed as NCHW is
    fine); scale/shift: optional fp32[3] device vectors of the per-channel normalisation v*scale + shift;
    padded_hw: the (H, W) >= (h, w) the network sees, the difference is zero padding (size divisor)."""
    n, c, h, w = x.shape
    op = _C.TdetOp()
    op.kind = _C.OP_PREP
    ph, pw = padded_hw if padded_hw is not None else (h, w)
    op.n, op.h, op.w, op.cin = n, ph, pw, c
    op.flags = _C.FLAG_SPLIT if split else 0   # y then holds 2n staged planes: hi then lo
    if (ph, pw) != (h, w):
        op.hc, op.wc = h, w
    op.ho, op.wo = ho, wo
    if x.dtype not in (torch.float32, torch.bfloat16, torch.uint8):
        raise NotImplementedError("input dtype %s (supported: float32, bfloat16, uint8)" % x.dtype)
    op.x_dtype = _TD[x.dtype]
    for i, s in enumerate(x.stride()):
        op.x_stride[i] = s
    op.x, op.y = _ptr(x), _ptr(y)
    op.y_dtype = _TD[y_dtype]   # format of the staging (bf16, or fp16: see TDET_OP_PREP)
    op.y_meta = y_meta
    op.scale, op.shift = _ptr(scale), _ptr(shift)
    return op


def op_stem(n, h, w, x, wgt, y, scale, shift, relu=True, x_meta=None, consts=None, scaled_out=False, split=False,
            pool=False, x_dtype=torch.bfloat16):
    """x: staged image buffer (x_dtype); y: ``Act`` of shape (n, ho, wo, 64) -- with ``pool`` the kernel also
    applies the 3x3/2 max-pool and y is the pooled (n, (ho-1)//2+1, (wo-1)//2+1, 64)."""
    op = _C.TdetOp()
    op.kind = _C.OP_STEM
    op.flags = (_C.FLAG_RELU if relu else 0) | (_C.FLAG_SCALED_OUT if scaled_out else 0) | \
        (_C.FLAG_SPLIT if split else 0) | (_C.FLAG_POOL if pool else 0)
    op.n, op.h, op.w, op.cin = n, h, w, 3
    op.cout, op.kh, op.kw = 64, 7, 7
    op.stride, op.pad, op.dil = 2, 3, 1
    op.ho, op.wo = conv_out(h, 7, 2, 3), conv_out(w, 7, 2, 3)
    op.x_dtype, op.y_dtype = _TD[x_dtype], _TD[y.dtype]
    op.x, op.wgt, op.y = _ptr(x), _ptr(wgt), y.ptr
    op.x_meta, op.y_meta = x_meta, y.meta
    op.scale, op.shift = _ptr(scale), _ptr(shift)
    op.bound_consts = _ptr(consts)
    return op


def op_maxpool(x, y, split=False):
    n, h, w, c = x.shape
    op = _C.TdetOp()
    op.kind = _C.OP_MAXPOOL
    op.flags = _C.FLAG_SPLIT if split else 0
    op.n, op.h, op.w, op.cin = n, h, w, c
    op.cout, op.kh, op.kw, op.stride, op.pad, op.dil = c, 3, 3, 2, 1, 1
    op.ho, op.wo = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)
    op.x_dtype = op.y_dtype = _TD[x.dtype]
    op.x, op.y = x.ptr, y.ptr
    return op


def op_split_combine(x, y):
    """y (fp32 NHWC tensor) = hi + lo of the split-precision activation x (logical shape (n, h, w, c))."""
    n, h, w, c = x.shape
    assert y.dtype == torch.float32 and y.numel() == n * h * w * c
    op = _C.TdetOp()
    op.kind = _C.OP_SPLIT_COMBINE
    op.n, op.h, op.w, op.cin = n, h, w, c
    op.x, op.y = x.ptr, y.data_ptr()
    return op


def op_subsample(x, y, split=False):
    n, h, w, c = x.shape
    if split:
        c = 2 * c   # a pure copy: the physical channel count is what matters
    op = _C.TdetOp()
    op.kind = _C.OP_SUBSAMPLE
    op.n, op.h, op.w, op.cin = n, h, w, c
    op.cout, op.kh, op.kw, op.stride, op.pad, op.dil = c, 1, 1, 2, 0, 1
    op.ho, op.wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    op.x_dtype = op.y_dtype = _TD[x.dtype]
    op.x, op.y = x.ptr, y.ptr
    return op


def _op_geom(op, x, cout, kh, kw, stride, pad, dil):
    n, h, w, cin = x.shape
    op.n, op.h, op.w, op.cin = n, h, w, cin
    op.cout, op.kh, op.kw = cout, kh, kw
    op.stride, op.pad, op.dil = stride, pad, dil
    op.ho, op.wo = conv_out(h, kh, stride, pad, dil), conv_out(w, kw, stride, pad, dil)


def op_wgrad(x, gy, dw, kh, kw, stride, pad, dil=1, scale=None):
    """x: forward input ``Act``; gy: output-gradient ``Act`` (bf16); dw: fp32 tensor [cout][kh][kw][cin],
    accumulated."""
    op = _C.TdetOp()
    op.kind = _C.OP_WGRAD
    _op_geom(op, x, gy.shape[3], kh, kw, stride, pad, dil)
    assert gy.shape == (op.n, op.ho, op.wo, op.cout), (gy.shape, (op.n, op.ho, op.wo, op.cout))
    assert dw.dtype == torch.float32 and dw.numel() == op.cout * kh * kw * op.cin
    op.x_dtype, op.gy_dtype = _TD[x.dtype], _TD[gy.dtype]
    op.x, op.gy, op.dw = x.ptr, gy.ptr, dw.data_ptr()
    op.x_meta, op.gy_meta = x.meta, gy.meta
    op.scale = _ptr(scale)
    return op


def op_dw_unpack(src, dst, cout, cin, kh, kw, groups=1):
    """fp32 [cout][kh][kw][cin] accumulator -> fp32 OIHW gradient [cout][cin / groups][kh][kw] (groups > 1: the block
    diagonal of the dense gradient)."""
    op = _C.TdetOp()
    op.kind = _C.OP_DW_UNPACK
    op.groups = groups
    op.cout, op.cin, op.kh, op.kw = cout, cin, kh, kw
    op.x, op.y = src.data_ptr(), dst.data_ptr()
    return op


def op_colsum(x, dw):
    n, h, w, c = x.shape
    op = _C.TdetOp()
    op.kind = _C.OP_COLSUM
    op.n, op.h, op.w, op.cin = n, h, w, c
    op.x_dtype = _TD[x.dtype]
    op.x, op.dw = x.ptr, dw.data_ptr()
    return op


def _op_eltwise(kind, x, y):
    n, h, w, c = x.shape
    op = _C.TdetOp()
    op.kind = kind
    op.n, op.h, op.w, op.cin = n, h, w, c
    op.cout = c
    op.ho, op.wo = y.shape[1], y.shape[2]
    op.x_dtype, op.y_dtype = _TD[x.dtype], _TD[y.dtype]
    op.x, op.y = x.ptr, y.ptr
    return op


def op_sumpool2(x, y):
    return _op_eltwise(_C.OP_SUMPOOL2, x, y)


def op_dilate2(x, y):
    return _op_eltwise(_C.OP_DILATE2, x, y)


def op_add_mask(x, y, residual=None, mask=None, scaled_out=False):
    """y = (x + residual) * (mask != 0) across 16-bit formats / per-tensor exponents; scaled_out: y (fp16)
    gets an exponent chosen from the inputs' recorded |max| (their metas must carry it)."""
    op = _op_eltwise(_C.OP_ADD_MASK, x, y)
    op.x_meta, op.y_meta = x.meta, y.meta
    if scaled_out:
        op.flags = _C.FLAG_SCALED_OUT
    if residual is not None:
        op.residual, op.residual_dtype, op.residual_meta = residual.ptr, _TD[residual.dtype], residual.meta
    if mask is not None:
        op.mask = mask.ptr
    return op


def op_parity_merge(parts, y, hc, wc, mask=None, scaled_out=False):
    """y[2i+a][2j+b] = parts[2a+b][i+o][j+o] * (mask != 0): interleaves the four parity-class results of a stride-2
    3x3 dgrad (TDET_OP_PARITY_MERGE; o = 1 for the three pad-1 classes).  hc, wc = size of the coarse gradient."""
    n, h, w, c = y.shape
    p00, p01, p10, p11 = parts
    assert p00.shape == (n, hc, wc, c) and p01.shape == (n, hc + 2, wc + 1, c)
    assert p10.shape == (n, hc + 1, wc + 2, c) and p11.shape == (n, hc + 1, wc + 1, c)
    op = _C.TdetOp()
    op.kind = _C.OP_PARITY_MERGE
    op.n, op.h, op.w, op.cin, op.hc, op.wc = n, h, w, c, hc, wc
    op.x_dtype = op.residual_dtype = op.coarse_dtype = op.gy_dtype = _TD[p00.dtype]
    op.y_dtype = _TD[y.dtype]
    op.x, op.residual, op.coarse, op.gy = p00.ptr, p01.ptr, p10.ptr, p11.ptr
    op.x_meta, op.residual_meta, op.coarse_meta, op.gy_meta = p00.meta, p01.meta, p10.meta, p11.meta
    op.y, op.y_meta = y.ptr, y.meta
    if mask is not None:
        op.mask = mask.ptr
    if scaled_out:
        op.flags = _C.FLAG_SCALED_OUT
    return op


def op_bn_affine_grad(g, a, gamma, beta, dw, b=None):
    """dw[0:C] += dgamma, dw[C:2C] += dbeta of a frozen-statistics BatchNorm: g = masked gradient w.r.t. its
    output, a (minus b, if given) = the stored tensor that equals the BN output wherever g != 0."""
    n, h, w, c = g.shape
    assert a.shape == g.shape and dw.numel() == 2 * c and dw.dtype == torch.float32
    op = _C.TdetOp()
    op.kind = _C.OP_BN_AFFINE_GRAD
    op.n, op.h, op.w, op.cin = n, h, w, c
    op.gy, op.gy_dtype, op.gy_meta = g.ptr, _TD[g.dtype], g.meta
    op.x, op.x_dtype, op.x_meta = a.ptr, _TD[a.dtype], a.meta
    if b is not None:
        op.residual, op.residual_dtype, op.residual_meta = b.ptr, _TD[b.dtype], b.meta
    op.scale, op.shift = gamma.data_ptr(), beta.data_ptr()
    op.dw = dw.data_ptr()
    return op


def op_maxpool_bwd(s, g, ds):
    """ds (bf16) = (s > 0) * scatter of g to each 3x3/2 window's first maximum of s (max-pool + ReLU backward)."""
    n, h, w, c = s.shape
    op = _C.TdetOp()
    op.kind = _C.OP_MAXPOOL_BWD
    op.n, op.h, op.w, op.cin = n, h, w, c
    op.ho, op.wo = g.shape[1], g.shape[2]
    op.x, op.x_dtype = s.ptr, _TD[s.dtype]
    op.gy, op.gy_dtype, op.gy_meta = g.ptr, _TD[g.dtype], g.meta
    op.y, op.y_dtype = ds.ptr, _C.BF16
    return op


def op_stem_wgrad(n, h, w, staged, g, dw, scale=None):
    """dw (fp32 [64][3][7][7]) += scale * wgrad of the stem conv; staged = TDET_OP_PREP output, g bf16 (n, ho, wo, 64)."""
    op = _C.TdetOp()
    op.kind = _C.OP_STEM_WGRAD
    op.n, op.h, op.w, op.cin, op.cout = n, h, w, 3, 64
    op.ho, op.wo = g.shape[1], g.shape[2]
    op.x, op.gy, op.gy_dtype = _ptr(staged), g.ptr, _TD[g.dtype]
    op.scale = _ptr(scale)
    op.dw = dw.data_ptr()
    return op


def gn_stats_numel(n, groups):
    """Size of the statistics buffer of one GroupNorm: fp32 [n][TDET_GN_STAT_BLOCKS][groups][2]."""
    return n * _C.GN_STAT_BLOCKS * groups * 2


def op_gn_stats(x, stats, groups):
    """stats (fp32 [n][TDET_GN_STAT_BLOCKS][groups][2]) <- per-block partial sum / sum of squares of x's true values
    per (image, group); no atomics, bit-reproducible."""
    n, h, w, c = x.shape
    assert stats.dtype == torch.float32 and stats.numel() == gn_stats_numel(n, groups)
    op = _C.TdetOp()
    op.kind = _C.OP_GN_STATS
    op.n, op.h, op.w, op.cin, op.groups = n, h, w, c, groups
    op.x, op.x_dtype, op.x_meta = x.ptr, _TD[x.dtype], x.meta
    op.dw = stats.data_ptr()
    return op


def op_gn_apply(x, stats, groups, gamma, beta, eps, y, residual=None, coarse=None, relu=False, relu6=False):
    """y = act(GroupNorm(x) + residual + up2(coarse)) from the statistics of op_gn_stats; y has exponent 0 and its
    meta (if any) receives max |y|."""
    n, h, w, c = x.shape
    assert y.shape == x.shape
    op = _C.TdetOp()
    op.kind = _C.OP_GN_APPLY
    op.flags = _C.FLAG_RELU6 if relu6 else (_C.FLAG_RELU if relu else 0)
    op.n, op.h, op.w, op.cin, op.groups = n, h, w, c, groups
    op.ho, op.wo, op.cout = h, w, c
    op.x, op.x_dtype, op.x_meta = x.ptr, _TD[x.dtype], x.meta
    op.y, op.y_dtype, op.y_meta = y.ptr, _TD[y.dtype], y.meta
    op.dw = stats.data_ptr()
    op.scale, op.shift = gamma.data_ptr(), beta.data_ptr()
    op.eps = eps
    if residual is not None:
        assert residual.shape == x.shape
        op.residual, op.residual_dtype, op.residual_meta = residual.ptr, _TD[residual.dtype], residual.meta
    if coarse is not None:
        op.coarse, op.coarse_dtype, op.coarse_meta = coarse.ptr, _TD[coarse.dtype], coarse.meta
        op.hc, op.wc = coarse.shape[1], coarse.shape[2]
    return op


def op_amax(x, meta):
    """meta.amax = max |x| (true values); `meta` = device address of a tdet_tensor_meta."""
    n, h, w, c = x.shape
    op = _C.TdetOp()
    op.kind = _C.OP_AMAX
    op.n, op.h, op.w, op.cin = n, h, w, c
    op.x_dtype = _TD[x.dtype]
    op.x, op.x_meta, op.y_meta = x.ptr, x.meta, meta
    return op


def op_zero(t):
    op = _C.TdetOp()
    op.kind = _C.OP_ZERO
    op.y = t.data_ptr()
    op.x_stride[0] = t.numel() * t.element_size()
    return op


def run_op(op, device):
    """Runs one op immediately on the current stream of `device` (tests / debugging)."""
    idx = _index(device)
    with torch.cuda.device(idx):
        _C.check(_C.lib().tdet_op_run(ctypes.byref(op), idx, _stream_ptr(device)))


def debug_im2col_tile(op, m0, r, s, kc, device):
    idx = _index(device)
    out = torch.empty((128, 64), dtype=torch.bfloat16, device=device)
    with torch.cuda.device(idx):
        _C.check(_C.lib().tdet_debug_im2col_tile(ctypes.byref(op), m0, r, s, kc, out.data_ptr(), idx,
                                                 _stream_ptr(device)))
    return out


# --------------------------------------------------------------------------------------------
# plans
# --------------------------------------------------------------------------------------------

class Plan:
    """A validated op sequence with prebuilt TMA descriptors (``tdet_plan``).

    `keepalive` holds every tensor the ops point at (packed weights, folded BN vectors, workspace
    activations) so the device memory outlives the plan.  `ext` are the tensors whose pointers may be
    re-bound per run (network inputs and returned outputs).  `meta` is the plan's ``MetaArena``."""

    def __init__(self, ops, ext, keepalive, device, meta=None):
        self.device = device
        self.index = _index(device)
        self.keepalive = keepalive
        self.meta = meta
        self.n_ext = len(ext)
        arr = (_C.TdetOp * len(ops))(*ops)
        ext_arr = (ctypes.c_void_p * max(1, self.n_ext))(*[t.data_ptr() for t in ext])
        ext_bytes = (ctypes.c_size_t * max(1, self.n_ext))(*[t.numel() * t.element_size() for t in ext])
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.index):
            _C.check(_C.lib().tdet_plan_create(
                ctypes.byref(handle), arr, len(ops), ext_arr, ext_bytes, self.n_ext,
                meta.tensor.data_ptr() if meta is not None else None,
                meta.count if meta is not None else 0, self.index))
        self._handle = handle
        self.num_launches = _C.lib().tdet_plan_num_launches(handle)
        self.flops = _C.lib().tdet_plan_flops(handle)

    def _ext_array(self, ext):
        assert len(ext) == self.n_ext
        return (ctypes.c_void_p * max(1, self.n_ext))(*[t.data_ptr() for t in ext])

    def run(self, ext):
        with torch.cuda.device(self.index):
            _C.check(_C.lib().tdet_plan_run(self._handle, self._ext_array(ext), self.n_ext,
                                            _stream_ptr(self.device)))

    def run_range(self, ext, first, last):
        """Runs ops [first, last) only, so the caller can interleave stream work between segments."""
        with torch.cuda.device(self.index):
            _C.check(_C.lib().tdet_plan_run_range(self._handle, self._ext_array(ext), self.n_ext, first,
                                                  last, _stream_ptr(self.device)))

    def run_timed(self, ext):
        """Per-launch device milliseconds (CUDA events between launches); measurement only."""
        ms = (ctypes.c_float * self.num_launches)()
        with torch.cuda.device(self.index):
            _C.check(_C.lib().tdet_plan_run_timed(self._handle, self._ext_array(ext), self.n_ext,
                                                  _stream_ptr(self.device), ms))
        return list(ms)

    def launch_info(self):
        out = []
        for i in range(self.num_launches):
            info = _C.TdetLaunchInfo()
            _C.check(_C.lib().tdet_plan_launch_info(self._handle, i, ctypes.byref(info)))
            out.append({f: getattr(info, f) for f, _ in _C.TdetLaunchInfo._fields_})
        return out

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _C.lib().tdet_plan_destroy(h)
            except Exception:
                pass
            self._handle = None
