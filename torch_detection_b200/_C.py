"""ctypes binding of ``libtdet_b200.so`` (C ABI: ``include/tdet_b200.h``).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception
is raised.  Build it with ``python -m torch_detection_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtdet_b200.so")
ABI_VERSION = 11

# tdet_status
OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_UNSUPPORTED_SHAPE = -2
ERR_UNSUPPORTED_DEVICE = -3
ERR_CUDA = -4
ERR_DRIVER = -5
ERR_OUT_OF_MEMORY = -6

# tdet_op_kind
OP_PREP, OP_STEM, OP_MAXPOOL, OP_CONV, OP_SUBSAMPLE = 0, 1, 2, 3, 4
OP_WGRAD, OP_DW_UNPACK, OP_COLSUM, OP_SUMPOOL2, OP_DILATE2, OP_ADD_MASK, OP_ZERO = 5, 6, 7, 8, 9, 10, 11
OP_AMAX = 12
OP_BN_AFFINE_GRAD = 13
OP_SPLIT_COMBINE = 14
OP_MAXPOOL_BWD, OP_STEM_WGRAD, OP_PARITY_MERGE = 15, 16, 17
OP_BOTTLENECK_TAIL = 18
OP_GN_STATS, OP_GN_APPLY = 19, 20
GN_STAT_BLOCKS = 128  # TDET_GN_STAT_BLOCKS
# tdet_dtype
BF16, F32, F16, U8 = 0, 1, 2, 3
FLAG_RELU = 1
FLAG_SCALED_OUT = 2
FLAG_COARSE_PARITY = 4
FLAG_SPLIT = 8
FLAG_POOL = 16
FLAG_RELU6 = 128
FLAG_DUAL = 32
FLAG_SCALED_OUT2 = 64
FLAG_REVERSE = 256


class TdetOp(ctypes.Structure):
    """Mirror of ``struct tdet_op``."""
    _fields_ = [
        ("kind", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("n", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32), ("cin", ctypes.c_int32),
        ("cout", ctypes.c_int32), ("kh", ctypes.c_int32), ("kw", ctypes.c_int32),
        ("stride", ctypes.c_int32), ("pad", ctypes.c_int32), ("dil", ctypes.c_int32),
        ("ho", ctypes.c_int32), ("wo", ctypes.c_int32),
        ("hc", ctypes.c_int32), ("wc", ctypes.c_int32),
        ("x_dtype", ctypes.c_int32), ("y_dtype", ctypes.c_int32),
        ("residual_dtype", ctypes.c_int32), ("coarse_dtype", ctypes.c_int32),
        ("x_stride", ctypes.c_int64 * 4),
        ("x", ctypes.c_void_p), ("wgt", ctypes.c_void_p), ("y", ctypes.c_void_p),
        ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p),
        ("residual", ctypes.c_void_p), ("coarse", ctypes.c_void_p),
        ("x_meta", ctypes.c_void_p), ("residual_meta", ctypes.c_void_p),
        ("coarse_meta", ctypes.c_void_p), ("y_meta", ctypes.c_void_p),
        ("bound_consts", ctypes.c_void_p),
        ("mask", ctypes.c_void_p), ("gy", ctypes.c_void_p),
        ("gy_dtype", ctypes.c_int32), ("groups", ctypes.c_int32),
        ("dw", ctypes.c_void_p), ("gy_meta", ctypes.c_void_p),
        ("x2", ctypes.c_void_p), ("x2_meta", ctypes.c_void_p),
        ("cin2", ctypes.c_int32), ("stride2", ctypes.c_int32), ("h2", ctypes.c_int32), ("w2", ctypes.c_int32),
        ("x2_dtype", ctypes.c_int32), ("eps", ctypes.c_float),
        ("wgt2", ctypes.c_void_p), ("scale2", ctypes.c_void_p), ("shift2", ctypes.c_void_p),
        ("bound_consts2", ctypes.c_void_p),
        ("wgt3", ctypes.c_void_p), ("scale3", ctypes.c_void_p), ("shift3", ctypes.c_void_p),
        ("bound_consts3", ctypes.c_void_p),
        ("y2", ctypes.c_void_p), ("y2_meta", ctypes.c_void_p),
        ("cout2", ctypes.c_int32), ("cout3", ctypes.c_int32), ("y2_dtype", ctypes.c_int32), ("reserved1", ctypes.c_int32),
    ]


class TdetLaunchInfo(ctypes.Structure):
    """Mirror of ``struct tdet_launch_info``."""
    _fields_ = [("kind", ctypes.c_int32), ("tile_n", ctypes.c_int32), ("grid", ctypes.c_int32),
                ("a_mode", ctypes.c_int32), ("m", ctypes.c_int32), ("n", ctypes.c_int32),
                ("k", ctypes.c_int32), ("variant", ctypes.c_int32),
                ("flops", ctypes.c_double), ("bytes", ctypes.c_double)]


class TdetError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libtdet_b200 error %d: %s" % (code, msg))
        self.code = code


EXPORTS = [
    "tdet_abi_version", "tdet_last_error", "tdet_device_supported",
    "tdet_stem_staging_dims", "tdet_set_sm_reserve",
    "tdet_pack_conv_weight", "tdet_pack_conv_weight_scaled", "tdet_pack_conv_weight_split", "tdet_pack_stem_weight_split",
    "tdet_pack_grouped_conv_weight", "tdet_pack_dgrad_weight", "tdet_pack_stem_weight", "tdet_fold_bn",
    "tdet_conv_bound_consts",
    "tdet_op_run", "tdet_plan_create", "tdet_plan_run", "tdet_plan_run_range", "tdet_plan_run_timed",
    "tdet_plan_num_launches", "tdet_plan_launch_info",
    "tdet_plan_flops", "tdet_plan_destroy", "tdet_debug_im2col_tile",
]

_lib = None


def lib():
    """Loads the extension (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libtdet_b200.so is not built (%s missing). Run `python -m torch_detection_b200.build`. "
            "There is no CPU/PyTorch fallback for this path." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    L.tdet_abi_version.restype = i32
    L.tdet_last_error.restype = ctypes.c_char_p
    L.tdet_device_supported.argtypes = [i32]
    L.tdet_set_sm_reserve.argtypes = [i32, i32]
    L.tdet_pack_conv_weight.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    L.tdet_pack_conv_weight_scaled.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    L.tdet_pack_conv_weight_split.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.tdet_pack_stem_weight_split.argtypes = [vp, vp, vp]
    L.tdet_pack_grouped_conv_weight.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp]
    L.tdet_pack_dgrad_weight.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp]
    L.tdet_pack_stem_weight.argtypes = [vp, vp, i32, vp]
    L.tdet_fold_bn.argtypes = [vp, vp, vp, vp, f32, vp, vp, i32, vp]
    L.tdet_conv_bound_consts.argtypes = [vp, i32, vp, vp, i32, i32, vp, vp]
    L.tdet_stem_staging_dims.argtypes = [i32, i32, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.tdet_op_run.argtypes = [ctypes.POINTER(TdetOp), i32, vp]
    L.tdet_plan_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(TdetOp), i32,
                                   ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t), i32, vp, i32,
                                   i32]
    L.tdet_plan_run.argtypes = [vp, ctypes.POINTER(vp), i32, vp]
    L.tdet_plan_run_range.argtypes = [vp, ctypes.POINTER(vp), i32, i32, i32, vp]
    L.tdet_plan_run_timed.argtypes = [vp, ctypes.POINTER(vp), i32, vp, ctypes.POINTER(f32)]
    L.tdet_plan_launch_info.argtypes = [vp, i32, ctypes.POINTER(TdetLaunchInfo)]
    L.tdet_plan_num_launches.argtypes = [vp]
    L.tdet_plan_flops.argtypes = [vp]
    L.tdet_plan_flops.restype = ctypes.c_double
    L.tdet_plan_destroy.argtypes = [vp]
    L.tdet_debug_im2col_tile.argtypes = [ctypes.POINTER(TdetOp), i32, i32, i32, i32, vp, i32, vp]
    for name in EXPORTS:
        if name not in ("tdet_last_error", "tdet_plan_flops"):
            getattr(L, name).restype = i32
    if L.tdet_abi_version() != ABI_VERSION:
        raise ImportError("libtdet_b200.so ABI version mismatch (rebuild it)")
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise TdetError(rc, lib().tdet_last_error().decode("utf-8", "replace"))
