"""Tensor-level entry points used by the layers: marshal torch CUDA tensors (device path, zero copy) or NumPy / CPU
arrays (host path: the library stages host<->device itself) into the C ABI of include/qnn.h."""
import ctypes
import os

import numpy as np

from . import _native
from ._layer import conv_output_length


def default_math():
    return os.environ.get("QNN_MATH", "tf32").lower()


def default_algo():
    return os.environ.get("QNN_ALGO", "auto").lower()


def _is_torch(x):
    return type(x).__module__.split(".")[0] == "torch"


def _dev_ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _host_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _as_host_f32(x):
    if _is_torch(x):
        x = x.detach().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


class PackedKernels(object):
    """Per-layer cache of packed kernel images (include/qnn.h: qnn_*_pack): device buffers keyed by
    (kind, math, algo, device, problem signature), valid for one `Variable.version`.  The layer owns one; an image is
    rebuilt (one ~2 us launch) after every weight update and otherwise reused, so a steady-state forward is ONE launch."""

    MAX_ENTRIES = 16        # variable-length inputs may change the kernel family / image: keep a few, then start over

    def __init__(self):
        self.version = None
        self.images = {}

    def get(self, kernel, key, nbytes_fn, pack_fn, device):
        """Image for `key`, or None when the problem has no packed form (general kernel)."""
        import torch
        if self.version != kernel.version:
            self.images.clear()
            self.version = kernel.version
        hit = self.images.get(key)
        if hit is not None:
            return hit if hit is not False else None
        nbytes = int(nbytes_fn())
        if nbytes == 0:
            self.images[key] = False
            return None
        if len(self.images) >= self.MAX_ENTRIES:
            self.images.clear()
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        pack_fn(buf)
        self.images[key] = buf
        return buf


def conv_out_shape(in_shape, filters, kernel_size, strides, padding, data_format, dilation_rate):
    rank = len(kernel_size)
    space = in_shape[2:] if data_format == "channels_first" else in_shape[1:-1]
    new = tuple(conv_output_length(space[a], kernel_size[a], padding, strides[a], dilation_rate[a]) for a in range(rank))
    if data_format == "channels_first":
        return (in_shape[0], 4 * filters) + new
    return (in_shape[0],) + new + (4 * filters,)


def conv_forward(x, kernel, bias, filters, kernel_size, strides, padding, data_format, dilation_rate, activation,
                 math=None, algo=None, packed=None):
    """QuaternionConv.call (reference complexnn/conv.py:288-345) as one fused kernel launch.
    x: torch CUDA tensor (returns a torch CUDA tensor) or NumPy / torch CPU array (returns NumPy / torch CPU).
    kernel, bias: `Variable`s (bias may be None).  packed: the layer's `PackedKernels` cache (device path only)."""
    lib = _native.lib()
    rank = len(kernel_size)
    shape = tuple(int(s) for s in x.shape)
    if len(shape) != rank + 2:
        raise ValueError("expected ndim=%d, found ndim=%d" % (rank + 2, len(shape)))
    c_axis = 1 if data_format == "channels_first" else -1
    if shape[c_axis] % 4:
        raise ValueError("channel dimension %d is not a multiple of 4" % shape[c_axis])
    in_q = shape[c_axis] // 4
    if tuple(kernel.shape) != tuple(kernel_size) + (in_q, 4 * filters):
        raise ValueError("kernel shape %s does not match kernel_size + (in_q, 4*filters) = %s"
                         % (tuple(kernel.shape), tuple(kernel_size) + (in_q, 4 * filters)))
    space = shape[2:] if data_format == "channels_first" else shape[1:-1]
    desc = _native.make_conv_desc(rank, shape[0], space, in_q, filters, kernel_size, strides, dilation_rate, padding,
                                  data_format, activation, math or default_math(), algo or default_algo())
    out_shape = conv_out_shape(shape, filters, kernel_size, strides, padding, data_format, dilation_rate)
    out_shape = tuple(max(int(s), 0) for s in out_shape)
    if _is_torch(x) and x.is_cuda:
        import torch
        xt = x.detach()
        if xt.dtype != torch.float32 or not xt.is_contiguous():
            xt = xt.to(torch.float32).contiguous()
        y = torch.empty(out_shape, dtype=torch.float32, device=xt.device)
        with torch.cuda.device(xt.device):
            kd = kernel.device(xt.device)
            img = None
            if packed is not None and y.numel():
                key = ("conv", _native.PACK_FORWARD, desc.math, desc.algo, str(xt.device), tuple(space))
                img = packed.get(kernel, key, lambda: lib.qnn_conv_packed_bytes(ctypes.byref(desc), _native.PACK_FORWARD),
                                 lambda buf: _native.check(lib.qnn_conv_pack(ctypes.byref(desc), _native.PACK_FORWARD,
                                                                             _dev_ptr(kd), _dev_ptr(buf), _stream())),
                                 xt.device)
            bd = _dev_ptr(bias.device(xt.device)) if bias is not None else None
            if img is not None:
                _native.check(lib.qnn_conv_forward_packed(ctypes.byref(desc), _dev_ptr(xt), _dev_ptr(kd), _dev_ptr(img), bd,
                                                          _dev_ptr(y), _stream()))
            else:
                _native.check(lib.qnn_conv_forward(ctypes.byref(desc), _dev_ptr(xt), _dev_ptr(kd), bd, _dev_ptr(y),
                                                   _stream()))
        return y
    xh = _as_host_f32(x)
    y = np.empty(out_shape, dtype=np.float32)
    _native.check(lib.qnn_conv_forward_host(ctypes.byref(desc), _host_ptr(xh), _host_ptr(kernel.numpy()),
                                            _host_ptr(bias.numpy()) if bias is not None else None, _host_ptr(y), None))
    if _is_torch(x):
        import torch
        return torch.from_numpy(y)
    return y


def dense_forward(x, kernel, bias, units, activation, math=None, algo=None, packed=None):
    """QuaternionDense.call (reference complexnn/dense.py:126-164) as one fused kernel launch."""
    lib = _native.lib()
    shape = tuple(int(s) for s in x.shape)
    if len(shape) != 2:
        raise ValueError("expected ndim=2, found ndim=%d" % len(shape))
    rows, in_q, q_units = shape[0], shape[1] // 4, units // 4
    if tuple(kernel.shape) != (in_q, units) or shape[1] != 4 * in_q or units != 4 * q_units:
        raise ValueError("kernel shape %s does not match (in_q, units) = %s" % (tuple(kernel.shape), (in_q, units)))
    m, a, act = _native.MATH[math or default_math()], _native.ALGO[algo or default_algo()], _native.ACT[activation]
    if _is_torch(x) and x.is_cuda:
        import torch
        xt = x.detach()
        if xt.dtype != torch.float32 or not xt.is_contiguous():
            xt = xt.to(torch.float32).contiguous()
        y = torch.empty((rows, units), dtype=torch.float32, device=xt.device)
        with torch.cuda.device(xt.device):
            kd = kernel.device(xt.device)
            img = None
            if packed is not None and rows:
                key = ("dense", _native.PACK_FORWARD, m, a, str(xt.device))
                img = packed.get(kernel, key,
                                 lambda: lib.qnn_dense_packed_bytes(rows, in_q, q_units, m, a, _native.PACK_FORWARD),
                                 lambda buf: _native.check(lib.qnn_dense_pack(rows, in_q, q_units, m, a, _native.PACK_FORWARD,
                                                                              _dev_ptr(kd), _dev_ptr(buf), _stream())),
                                 xt.device)
            bd = _dev_ptr(bias.device(xt.device)) if bias is not None else None
            if img is not None:
                _native.check(lib.qnn_dense_forward_packed(rows, in_q, q_units, _dev_ptr(xt), _dev_ptr(kd), _dev_ptr(img), bd,
                                                           act, m, a, _dev_ptr(y), _stream()))
            else:
                _native.check(lib.qnn_dense_forward(rows, in_q, q_units, _dev_ptr(xt), _dev_ptr(kd), bd, act, m, a,
                                                    _dev_ptr(y), _stream()))
        return y
    xh = _as_host_f32(x)
    y = np.empty((rows, units), dtype=np.float32)
    _native.check(lib.qnn_dense_forward_host(rows, in_q, q_units, _host_ptr(xh), _host_ptr(kernel.numpy()),
                                             _host_ptr(bias.numpy()) if bias is not None else None, act, m, a,
                                             _host_ptr(y), None))
    if _is_torch(x):
        import torch
        return torch.from_numpy(y)
    return y


def _grad_out(like_shape, out, device):
    import torch
    if out is None:
        return torch.empty(like_shape, dtype=torch.float32, device=device)
    if tuple(out.shape) != tuple(like_shape) or not out.is_contiguous() or out.dtype != torch.float32:
        raise ValueError("gradient output buffer must be a contiguous fp32 tensor of shape %s" % (tuple(like_shape),))
    return out


def conv_backward(x, y, dy, kernel, has_bias, filters, kernel_size, strides, padding, data_format, dilation_rate,
                  activation, need_dx=True, dkernel_out=None, dbias_out=None, math=None, algo=None, packed=None):
    """Gradients of the quaternion convolution (device tensors only).  Returns (dx | None, dkernel, dbias | None);
    dkernel_out / dbias_out may be views into a flat gradient bucket."""
    import torch
    lib = _native.lib()
    rank = len(kernel_size)
    shape = tuple(int(s) for s in x.shape)
    c_axis = 1 if data_format == "channels_first" else -1
    in_q = shape[c_axis] // 4
    space = shape[2:] if data_format == "channels_first" else shape[1:-1]
    desc = _native.make_conv_desc(rank, shape[0], space, in_q, filters, kernel_size, strides, dilation_rate, padding,
                                  data_format, activation, math or default_math(), algo or default_algo())
    x, y, dy = x.contiguous(), y.contiguous(), dy.contiguous()
    dx = torch.empty_like(x) if need_dx else None
    dk = _grad_out(tuple(kernel.shape), dkernel_out, x.device)
    db = _grad_out((4 * filters,), dbias_out, x.device) if has_bias else None
    with torch.cuda.device(x.device):
        kd = kernel.device(x.device)
        img = None
        if packed is not None and need_dx and dy.numel():
            key = ("conv", _native.PACK_DGRAD, desc.math, desc.algo, str(x.device), tuple(space))
            img = packed.get(kernel, key, lambda: lib.qnn_conv_packed_bytes(ctypes.byref(desc), _native.PACK_DGRAD),
                             lambda buf: _native.check(lib.qnn_conv_pack(ctypes.byref(desc), _native.PACK_DGRAD, _dev_ptr(kd),
                                                                         _dev_ptr(buf), _stream())), x.device)
        _native.check(lib.qnn_conv_backward_packed(ctypes.byref(desc), _dev_ptr(x), _dev_ptr(kd), _dev_ptr(img), _dev_ptr(y),
                                                   _dev_ptr(dy), _dev_ptr(dx), _dev_ptr(dk), _dev_ptr(db), _stream()))
    return dx, dk, db


def dense_backward(x, y, dy, kernel, has_bias, units, activation, need_dx=True, dkernel_out=None, dbias_out=None,
                   math=None, algo=None, packed=None):
    import torch
    lib = _native.lib()
    rows, in_q, q_units = int(x.shape[0]), int(x.shape[1]) // 4, units // 4
    x, y, dy = x.contiguous(), y.contiguous(), dy.contiguous()
    dx = torch.empty_like(x) if need_dx else None
    dk = _grad_out(tuple(kernel.shape), dkernel_out, x.device)
    db = _grad_out((units,), dbias_out, x.device) if has_bias else None
    m, a = _native.MATH[math or default_math()], _native.ALGO[algo or default_algo()]
    with torch.cuda.device(x.device):
        kd = kernel.device(x.device)
        img = None
        if packed is not None and need_dx and rows:
            key = ("dense", _native.PACK_DGRAD, m, a, str(x.device))
            img = packed.get(kernel, key, lambda: lib.qnn_dense_packed_bytes(rows, in_q, q_units, m, a, _native.PACK_DGRAD),
                             lambda buf: _native.check(lib.qnn_dense_pack(rows, in_q, q_units, m, a, _native.PACK_DGRAD,
                                                                          _dev_ptr(kd), _dev_ptr(buf), _stream())), x.device)
        _native.check(lib.qnn_dense_backward_packed(rows, in_q, q_units, _dev_ptr(x), _dev_ptr(kd), _dev_ptr(img), _dev_ptr(y),
                                                    _dev_ptr(dy), _native.ACT[activation], m, a, _dev_ptr(dx), _dev_ptr(dk),
                                                    _dev_ptr(db), _stream()))
    return dx, dk, db
