"""Host-side glue between torch tensors (device memory, streams) and the C ABI.

torch is plumbing here: it owns device memory and the current stream.  Every arithmetic step of the
path runs in ``libtdet_b200.so``; nothing in this module computes on tensors with torch ops.
"""
import ctypes

import torch

from . import _C

BN_EPS = 1e-5  # nn.BatchNorm2d default (reference models/utils/layers.py:50-54)


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dev_index(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def require_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise NotImplementedError(
            "%s must be a CUDA tensor: the B200 path has no CPU fallback" % what)


def weight_dtype():
    """torch dtype of packed conv weights (what tdet_weight_dtype() reports)."""
    return torch.float16 if _C.lib().tdet_weight_dtype() == _C.F16 else torch.bfloat16


def conv_out(v, k, s, p, d=1):
    return (v + 2 * p - d * (k - 1) - 1) // s + 1


# --------------------------------------------------------------------------------------------
# operand preparation
# --------------------------------------------------------------------------------------------

def pack_conv_weight(w):
    """fp32 OIHW parameter -> bf16 [O][kh][kw][I] (tcgen05 B operand rows)."""
    require_cuda(w, "weight")
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    o, i, kh, kw = w.shape
    out = torch.empty((o, kh, kw, i), dtype=weight_dtype(), device=w.device)
    with torch.cuda.device(w.device):
        _C.check(_C.lib().tdet_pack_conv_weight(w.data_ptr(), out.data_ptr(), o, i, kh, kw,
                                                _stream_ptr(w.device)))
    return out


def pack_stem_weight(w):
    """fp32 [64][3][7][7] -> bf16 [64][448] stem operand."""
    require_cuda(w, "weight")
    w = w.detach()
    if tuple(w.shape) != (64, 3, 7, 7):
        raise ValueError("stem weight must be (64,3,7,7)")
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    out = torch.empty((64, 448), dtype=weight_dtype(), device=w.device)
    with torch.cuda.device(w.device):
        _C.check(_C.lib().tdet_pack_stem_weight(w.data_ptr(), out.data_ptr(), _stream_ptr(w.device)))
    return out


def fold_bn(bn):
    """eval-mode nn.BatchNorm2d -> (scale, shift) fp32 device vectors."""
    g, b, m, v = (t.detach().float().contiguous() for t in
                  (bn.weight, bn.bias, bn.running_mean, bn.running_var))
    require_cuda(g, "BatchNorm parameters")
    ch = g.numel()
    scale = torch.empty(ch, dtype=torch.float32, device=g.device)
    shift = torch.empty(ch, dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        _C.check(_C.lib().tdet_fold_bn(g.data_ptr(), b.data_ptr(), m.data_ptr(), v.data_ptr(),
                                       ctypes.c_float(bn.eps), scale.data_ptr(), shift.data_ptr(),
                                       ch, _stream_ptr(g.device)))
    return scale, shift


# --------------------------------------------------------------------------------------------
# op descriptors
# --------------------------------------------------------------------------------------------

def _ptr(t):
    return None if t is None else t.data_ptr()


def nhwc_empty(n, h, w, c, device):
    """Dense NHWC bf16 buffer exposed as a logical-NCHW channels_last tensor (zero-copy view)."""
    return torch.empty((n, c, h, w), dtype=torch.bfloat16, device=device,
                       memory_format=torch.channels_last)


def op_conv(x_shape, x, wgt, y, kh, kw, stride, pad, dil=1, scale=None, shift=None, residual=None,
            coarse=None, coarse_hw=(0, 0), relu=False):
    """x_shape = (n, h, w, cin) of the NHWC input; wgt packed [cout][kh][kw][cin]."""
    n, h, w, cin = x_shape
    cout = wgt.shape[0]
    op = _C.TdetOp()
    op.kind = _C.OP_CONV
    op.flags = _C.FLAG_RELU if relu else 0
    op.n, op.h, op.w, op.cin = n, h, w, cin
    op.cout, op.kh, op.kw = cout, kh, kw
    op.stride, op.pad, op.dil = stride, pad, dil
    op.ho, op.wo = conv_out(h, kh, stride, pad, dil), conv_out(w, kw, stride, pad, dil)
    op.hc, op.wc = coarse_hw
    op.x, op.wgt, op.y = _ptr(x), _ptr(wgt), _ptr(y)
    op.scale, op.shift = _ptr(scale), _ptr(shift)
    op.residual, op.coarse = _ptr(residual), _ptr(coarse)
    return op


def op_prep(x, y, ho, wo):
    n, c, h, w = x.shape
    op = _C.TdetOp()
    op.kind = _C.OP_PREP
    op.n, op.h, op.w, op.cin = n, h, w, c
    op.ho, op.wo = ho, wo
    if x.dtype == torch.float32:
        op.x_dtype = _C.F32
    elif x.dtype == torch.bfloat16:
        op.x_dtype = _C.BF16
    else:
        raise NotImplementedError("input dtype %s (supported: float32, bfloat16)" % x.dtype)
    for i, s in enumerate(x.stride()):
        op.x_stride[i] = s
    op.x, op.y = _ptr(x), _ptr(y)
    return op


def op_stem(n, h, w, x, wgt, y, scale, shift, relu=True):
    op = _C.TdetOp()
    op.kind = _C.OP_STEM
    op.flags = _C.FLAG_RELU if relu else 0
    op.n, op.h, op.w, op.cin = n, h, w, 3
    op.cout, op.kh, op.kw = 64, 7, 7
    op.stride, op.pad, op.dil = 2, 3, 1
    op.ho, op.wo = conv_out(h, 7, 2, 3), conv_out(w, 7, 2, 3)
    op.x, op.wgt, op.y = _ptr(x), _ptr(wgt), _ptr(y)
    op.scale, op.shift = _ptr(scale), _ptr(shift)
    return op


def op_maxpool(n, h, w, c, x, y):
    op = _C.TdetOp()
    op.kind = _C.OP_MAXPOOL
    op.n, op.h, op.w, op.cin = n, h, w, c
    op.cout, op.kh, op.kw, op.stride, op.pad, op.dil = c, 3, 3, 2, 1, 1
    op.ho, op.wo = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)
    op.x, op.y = _ptr(x), _ptr(y)
    return op


def op_subsample(n, h, w, c, x, y):
    op = _C.TdetOp()
    op.kind = _C.OP_SUBSAMPLE
    op.n, op.h, op.w, op.cin = n, h, w, c
    op.cout, op.kh, op.kw, op.stride, op.pad, op.dil = c, 1, 1, 2, 0, 1
    op.ho, op.wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    op.x, op.y = _ptr(x), _ptr(y)
    return op


def run_op(op, device):
    """Runs one op immediately on the current stream of `device` (tests / debugging)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    with torch.cuda.device(idx):
        _C.check(_C.lib().tdet_op_run(ctypes.byref(op), idx, _stream_ptr(device)))


def debug_im2col_tile(op, m0, r, s, kc, device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
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
    re-bound per run (network inputs and returned outputs)."""

    def __init__(self, ops, ext, keepalive, device):
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self.keepalive = keepalive
        self.n_ext = len(ext)
        arr = (_C.TdetOp * len(ops))(*ops)
        ext_arr = (ctypes.c_void_p * max(1, self.n_ext))(*[t.data_ptr() for t in ext])
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.index):
            _C.check(_C.lib().tdet_plan_create(ctypes.byref(handle), arr, len(ops), ext_arr,
                                               self.n_ext, self.index))
        self._handle = handle
        self.num_launches = _C.lib().tdet_plan_num_launches(handle)
        self.flops = _C.lib().tdet_plan_flops(handle)

    def run(self, ext):
        assert len(ext) == self.n_ext
        ext_arr = (ctypes.c_void_p * max(1, self.n_ext))(*[t.data_ptr() for t in ext])
        with torch.cuda.device(self.index):
            _C.check(_C.lib().tdet_plan_run(self._handle, ext_arr, self.n_ext,
                                            _stream_ptr(self.device)))

    def run_timed(self, ext):
        """Per-launch device milliseconds (CUDA events between launches); measurement only."""
        self.run(ext)
        ext_arr = (ctypes.c_void_p * max(1, self.n_ext))(*[t.data_ptr() for t in ext])
        ms = (ctypes.c_float * self.num_launches)()
        with torch.cuda.device(self.index):
            _C.check(_C.lib().tdet_plan_run_timed(self._handle, ext_arr, self.n_ext,
                                                  _stream_ptr(self.device), ms))
        return list(ms)

    def launch_info(self):
        out = []
        for i in range(self.num_launches):
            info = _C.TdetLaunchInfo()
            _C.check(_C.lib().tdet_plan_launch_info(self._handle, i, ctypes.byref(info)))
            out.append({f: getattr(info, f) for f, _ in _C.TdetLaunchInfo._fields_ if f != "reserved"})
        return out

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _C.lib().tdet_plan_destroy(h)
            except Exception:
                pass
            self._handle = None
