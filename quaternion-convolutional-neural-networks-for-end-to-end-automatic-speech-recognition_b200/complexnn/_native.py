"""ctypes binding of libqnn_b200.so (C ABI: include/qnn.h).  There is no CPU fallback: if the library is missing or a
call fails, an exception is raised."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# QNN_LIB_PATH: load another build of the same ABI (A/B timing of two builds on one box)
LIB_PATH = os.environ.get("QNN_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "lib", "libqnn_b200.so")

PAD = {"valid": 0, "same": 1, "causal": 2}
ACT = {None: 0, "linear": 0, "relu": 1, "tanh": 2, "sigmoid": 3, "hard_sigmoid": 4, "softplus": 5, "softsign": 6,
       "elu": 7, "selu": 8, "exponential": 9}
MATH = {"tf32": 0, "fp32": 1, "3xtf32": 2}
ALGO = {"auto": 0, "general": 1, "tensor": 2}

QNN_E_INVALID, QNN_E_UNSUPPORTED = -1, -2
ABI_VERSION = 2
PACK_FORWARD, PACK_DGRAD = 0, 1
KERNEL_GENERAL, KERNEL_TC_ROWS, KERNEL_TC_CF, KERNEL_SMALL_K = 0, 1, 2, 3


class ConvDesc(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_int32), ("batch", ctypes.c_int32), ("in_spatial", ctypes.c_int32 * 3),
                ("in_q", ctypes.c_int32), ("filters", ctypes.c_int32), ("kernel", ctypes.c_int32 * 3),
                ("stride", ctypes.c_int32 * 3), ("dilation", ctypes.c_int32 * 3), ("padding", ctypes.c_int32),
                ("channels_first", ctypes.c_int32), ("activation", ctypes.c_int32), ("math", ctypes.c_int32),
                ("algo", ctypes.c_int32)]


_lib = None
_P = ctypes.c_void_p

SIGNATURES = {
    "qnn_abi_version": (ctypes.c_int, []),
    "qnn_last_error": (ctypes.c_char_p, []),
    "qnn_launch_count": (ctypes.c_uint64, []),
    "qnn_conv_uses_tensor_cores": (ctypes.c_int, [ctypes.POINTER(ConvDesc)]),
    "qnn_conv_forward_kernel": (ctypes.c_int, [ctypes.POINTER(ConvDesc)]),
    "qnn_conv_work_split": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(ctypes.c_int32)]),
    "qnn_dense_forward_kernel": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                                ctypes.c_int32, ctypes.c_int32]),
    "qnn_dense_uses_tensor_cores": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]),
    "qnn_conv_backward_uses_tensor_cores": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(ctypes.c_int32),
                                                            ctypes.POINTER(ctypes.c_int32)]),
    "qnn_dense_backward_uses_tensor_cores": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                                             ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "qnn_conv_out_spatial": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(ctypes.c_int32 * 3)]),
    "qnn_conv_forward": (ctypes.c_int, [ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _P]),
    "qnn_dense_forward": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _P, _P, _P, ctypes.c_int32,
                                         ctypes.c_int32, ctypes.c_int32, _P, _P]),
    "qnn_conv_packed_bytes": (ctypes.c_size_t, [ctypes.POINTER(ConvDesc), ctypes.c_int32]),
    "qnn_conv_pack": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.c_int32, _P, _P, _P]),
    "qnn_conv_forward_packed": (ctypes.c_int, [ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P]),
    "qnn_dense_packed_bytes": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                                 ctypes.c_int32, ctypes.c_int32]),
    "qnn_dense_pack": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.c_int32, _P, _P, _P]),
    "qnn_dense_forward_packed": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _P, _P, _P, _P,
                                                ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P, _P]),
    "qnn_conv_backward_packed": (ctypes.c_int, [ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "qnn_dense_backward_packed": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _P, _P, _P, _P, _P,
                                                 ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P, _P, _P, _P]),
    "qnn_conv_backward": (ctypes.c_int, [ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P, _P]),
    "qnn_dense_backward": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _P, _P, _P, _P,
                                          ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P, _P, _P, _P]),
    "qnn_conv_forward_host": (ctypes.c_int, [ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _P]),
    "qnn_dense_forward_host": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _P, _P, _P,
                                              ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P, _P]),
    "qnn_debug_trace": (ctypes.c_int, [_P, ctypes.c_size_t]),
    "qnn_comm_unique_id": (ctypes.c_int, [_P]),
    "qnn_comm_init": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, _P]),
    "qnn_allreduce_f32": (ctypes.c_int, [_P, ctypes.c_size_t, _P]),
    "qnn_comm_destroy": (ctypes.c_int, []),
}


def lib():
    """The loaded library; raises RuntimeError when it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libqnn_b200.so is not built: run `python __graft_entry__.py` (or build.py) first; "
                               "expected at " + LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.qnn_abi_version() != ABI_VERSION:
            raise RuntimeError("libqnn_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc):
    """Translate a qnn_status into the exception type the reference would have raised at the same spot."""
    if rc == 0:
        return
    msg = lib().qnn_last_error().decode("utf-8", "replace")
    if rc == QNN_E_INVALID:
        raise ValueError(msg)
    if rc == QNN_E_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError("qnn error %d: %s" % (rc, msg))


def make_conv_desc(rank, batch, in_spatial, in_q, filters, kernel_size, strides, dilation_rate, padding, data_format,
                   activation, math="tf32", algo="auto"):
    d = ConvDesc()
    d.rank, d.batch, d.in_q, d.filters = rank, batch, in_q, filters
    for a in range(3):
        d.in_spatial[a] = in_spatial[a] if a < rank else 1
        d.kernel[a] = kernel_size[a] if a < rank else 1
        d.stride[a] = strides[a] if a < rank else 1
        d.dilation[a] = dilation_rate[a] if a < rank else 1
    d.padding = PAD[padding]
    d.channels_first = 1 if data_format == "channels_first" else 0
    d.activation = ACT[activation]
    d.math, d.algo = MATH[math], ALGO[algo]
    return d


def launch_count():
    return int(lib().qnn_launch_count())
