"""CPU oracle for the quaternion conv / dense hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
the shipped layers (the `complexnn` mirror inside the *_b200 package) never do and fail loudly without the CUDA library.

What it restates (reference = /root/reference, file:line):
  * Hamilton weight expansion of the convolution        complexnn/conv.py:294-331
  * the real convolution + bias + activation around it  complexnn/conv.py:309-343  (K.conv1d/2d/3d, K.bias_add)
  * Hamilton (transposed) expansion of the dense layer  complexnn/dense.py:131-143
  * matmul + bias + activation                          complexnn/dense.py:149-162
  * output-shape rule                                   complexnn/conv.py:347-372, complexnn/dense.py:166-171
  * weight initialisers (layout + RNG call order)       complexnn/init.py:48-93, 115-155
  * the gradients TF autodiff would produce through slice / negate / concat

Pinning.  The reference ships no tests or golden vectors and its arithmetic lives in Keras/TensorFlow, which are absent
from this image (setup.py:15-16 leaves them unpinned) -> PARITY AT THE KERAS/TF BOUNDARY IS UNPINNED.  What *is*
pinned: tests/golden/*.npz were produced by importing the reference's own, unmodified complexnn package on top of the
NumPy Keras stand-in in oracle/keras_shim (script: oracle/make_golden.py), so the slicing, sign table, concatenation
axes, stored-weight layout and initialiser RNG order below are checked against the reference's own code; the
third-party conv semantics (TF SAME/VALID/causal padding, no kernel flip) are cross-checked against
torch.nn.functional.conv{1,2,3}d in tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------------------------------------------
# Hamilton tables.  IDX[a][b]: which sub-filter (0=r,1=i,2=j,3=k) multiplies input component a
# into output component b; NEG[a][b]: 1 when that block enters with a minus sign.  Conv convention: y = W (x) x.
# The dense layer uses the transpose: IDX_D[a][b] = IDX[b][a], NEG_D[a][b] = NEG[b][a]  ->  y = conj(W) (x) x.
# ----------------------------------------------------------------------------------------------------------------------
IDX_CONV = np.array([[0, 1, 2, 3], [1, 0, 3, 2], [2, 3, 0, 1], [3, 2, 1, 0]])
NEG_CONV = np.array([[0, 0, 0, 0], [1, 0, 0, 1], [1, 1, 0, 0], [1, 0, 1, 0]])
IDX_DENSE = IDX_CONV.T.copy()
NEG_DENSE = NEG_CONV.T.copy()


def conv_output_length(n, k, padding, stride, dilation=1):
    """keras.utils.conv_utils.conv_output_length as used at complexnn/conv.py:352-358."""
    if n is None:
        return None
    eff = k + (k - 1) * (dilation - 1)
    if padding in ("same", "causal"):
        out = n
    elif padding == "valid":
        out = n - eff + 1
    else:
        raise ValueError("Invalid padding: %r" % (padding,))
    return (out + stride - 1) // stride


def pad_amounts(n, k, stride, dilation, padding):
    """(pad_before, pad_after, out_len) of tf.nn.convolution for one spatial axis."""
    eff = (k - 1) * dilation + 1
    if padding == "valid":
        return 0, 0, max((n - eff) // stride + 1, 0) if n >= eff else 0
    if padding == "same":
        out = -(-n // stride)
        total = max((out - 1) * stride + eff - n, 0)
        return total // 2, total - total // 2, out
    if padding == "causal":
        return dilation * (k - 1), 0, -(-n // stride)
    raise ValueError("Invalid padding: %r" % (padding,))


def _tup(v, rank):
    return (int(v),) * rank if np.isscalar(v) else tuple(int(e) for e in v)


# ----------------------------------------------------------------------------------------------------------------------
# Literal restatement: expand, then a stock real conv / matmul.
# ----------------------------------------------------------------------------------------------------------------------
def expand_conv_kernel(kernel, filters):
    """complexnn/conv.py:294-331.  kernel: kernel_size + (in_q, 4F)  ->  kernel_size + (4 in_q, 4F)."""
    F = filters
    f_r, f_i, f_j, f_k = (kernel[..., c * F:(c + 1) * F] for c in range(4))
    cat_r = np.concatenate([f_r, -f_i, -f_j, -f_k], axis=-2)
    cat_i = np.concatenate([f_i, f_r, -f_k, f_j], axis=-2)
    cat_j = np.concatenate([f_j, f_k, f_r, -f_i], axis=-2)
    cat_k = np.concatenate([f_k, -f_j, f_i, f_r], axis=-2)
    return np.concatenate([cat_r, cat_i, cat_j, cat_k], axis=-1)


def expand_dense_kernel(kernel, q_units):
    """complexnn/dense.py:131-143.  kernel: (in_q, 4q)  ->  (4 in_q, 4q)."""
    q = q_units
    r, i, j, k = (kernel[:, c * q:(c + 1) * q] for c in range(4))
    cat_r = np.concatenate([r, -i, -j, -k], axis=-1)
    cat_i = np.concatenate([i, r, -k, j], axis=-1)
    cat_j = np.concatenate([j, k, r, -i], axis=-1)
    cat_k = np.concatenate([k, -j, i, r], axis=-1)
    return np.concatenate([cat_r, cat_i, cat_j, cat_k], axis=0)


def real_conv(x, w, strides, padding, data_format, dilation, acc_dtype=np.float64):
    """K.conv1d/2d/3d (complexnn/conv.py:313-315,334): cross-correlation, kernel = spatial + (C_in, C_out)."""
    rank = w.ndim - 2
    strides, dilation = _tup(strides, rank), _tup(dilation, rank)
    if padding == "causal" and rank != 1:
        raise ValueError("causal padding is 1-D only")
    xs = np.moveaxis(x, 1, -1) if data_format == "channels_first" else x
    ksz = w.shape[:rank]
    pads, outs = [(0, 0)], []
    for a in range(rank):
        lo, hi, o = pad_amounts(xs.shape[1 + a], ksz[a], strides[a], dilation[a], padding)
        pads.append((lo, hi))
        outs.append(o)
    pads.append((0, 0))
    xp = np.pad(xs, pads).astype(acc_dtype, copy=False)
    wa = w.astype(acc_dtype, copy=False)
    y = np.zeros((xs.shape[0],) + tuple(outs) + (w.shape[-1],), dtype=acc_dtype)
    for tap in np.ndindex(*ksz):
        sl = [slice(None)]
        for a in range(rank):
            st = tap[a] * dilation[a]
            sl.append(slice(st, st + (outs[a] - 1) * strides[a] + 1, strides[a]))
        sl.append(slice(None))
        y += xp[tuple(sl)] @ wa[tap]
    return np.moveaxis(y, -1, 1) if data_format == "channels_first" else y


ACTIVATIONS = {
    None: lambda v: v,
    "linear": lambda v: v,
    "relu": lambda v: np.maximum(v, 0),
    "tanh": np.tanh,
    "sigmoid": lambda v: 1.0 / (1.0 + np.exp(-v)),
    "hard_sigmoid": lambda v: np.clip(0.2 * v + 0.5, 0.0, 1.0),
    "softplus": lambda v: np.logaddexp(v, 0.0),
    "softsign": lambda v: v / (1.0 + np.abs(v)),
    "elu": lambda v: np.where(v > 0, v, np.exp(np.minimum(v, 0)) - 1.0),
    "selu": lambda v: 1.0507009873554804934193349852946 * np.where(
        v > 0, v, 1.6732632423543772848170429916717 * (np.exp(np.minimum(v, 0)) - 1.0)),
    "exponential": np.exp,
}


def _softmax(v):
    e = np.exp(v - v.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


ACTIVATIONS["softmax"] = _softmax


def qconv_forward(x, kernel, bias, filters, strides=1, padding="valid", data_format="channels_last", dilation_rate=1,
                  activation=None, acc_dtype=np.float64, out_dtype=np.float32):
    """QuaternionConv.call, complexnn/conv.py:288-345 (rank inferred from the kernel)."""
    w_full = expand_conv_kernel(kernel, filters)
    y = real_conv(x, w_full, strides, padding, data_format, dilation_rate, acc_dtype)
    if bias is not None:
        b = bias.astype(acc_dtype)
        y = y + (b.reshape((1, -1) + (1,) * (y.ndim - 2)) if data_format == "channels_first" else b)
    y = ACTIVATIONS[activation](y)
    return y.astype(out_dtype) if out_dtype is not None else y


def qdense_forward(x, kernel, bias, units, activation=None, acc_dtype=np.float64, out_dtype=np.float32):
    """QuaternionDense.call, complexnn/dense.py:126-164 (the re-slice at :151-157 is a no-op)."""
    w_full = expand_dense_kernel(kernel, units // 4)
    y = x.astype(acc_dtype) @ w_full.astype(acc_dtype)
    if bias is not None:
        y = y + bias.astype(acc_dtype)
    y = ACTIVATIONS[activation](y)
    return y.astype(out_dtype) if out_dtype is not None else y


def qconv_abs_bound(x, kernel, filters, strides=1, padding="valid", data_format="channels_last", dilation_rate=1):
    """sum_k |x_k| |w_k| of every output's dot product (no bias): the scale a componentwise rounding bound refers to.
    A product of two operands each rounded to nearest tf32 (10 mantissa bits) is off by at most 2^-10 |x_k w_k|."""
    w_abs = np.abs(expand_conv_kernel(kernel, filters))
    return real_conv(np.abs(x), w_abs, strides, padding, data_format, dilation_rate)


def qdense_abs_bound(x, kernel, units):
    return np.abs(x).astype(np.float64) @ np.abs(expand_dense_kernel(kernel, units // 4)).astype(np.float64)


def qconv_output_shape(input_shape, filters, kernel_size, strides, padding, data_format, dilation_rate):
    """complexnn/conv.py:347-372."""
    rank = len(kernel_size)
    strides, dilation_rate = _tup(strides, rank), _tup(dilation_rate, rank)
    space = input_shape[1:-1] if data_format == "channels_last" else input_shape[2:]
    new = tuple(conv_output_length(space[a], kernel_size[a], padding, strides[a], dilation_rate[a]) for a in range(rank))
    if data_format == "channels_last":
        return (input_shape[0],) + new + (4 * filters,)
    return (input_shape[0], 4 * filters) + new


# ----------------------------------------------------------------------------------------------------------------------
# Independent formulation: 16 signed component-block products, no expanded weight (checks the literal one above).
# ----------------------------------------------------------------------------------------------------------------------
def qconv_forward_direct(x, kernel, bias, filters, strides=1, padding="valid", data_format="channels_last",
                         dilation_rate=1, activation=None):
    F = filters
    ch_axis = 1 if data_format == "channels_first" else -1
    C = x.shape[ch_axis] // 4
    xs = [np.take(x, np.arange(a * C, (a + 1) * C), axis=ch_axis) for a in range(4)]
    fs = [kernel[..., c * F:(c + 1) * F] for c in range(4)]
    outs = []
    for b in range(4):
        acc = 0
        for a in range(4):
            t = real_conv(xs[a], fs[IDX_CONV[a][b]], strides, padding, data_format, dilation_rate)
            acc = acc - t if NEG_CONV[a][b] else acc + t
        outs.append(acc)
    y = np.concatenate(outs, axis=ch_axis)
    if bias is not None:
        b64 = bias.astype(np.float64)
        y = y + (b64.reshape((1, -1) + (1,) * (y.ndim - 2)) if data_format == "channels_first" else b64)
    return ACTIVATIONS[activation](y).astype(np.float32)


def qdense_forward_direct(x, kernel, bias, units, activation=None):
    q = units // 4
    C = x.shape[-1] // 4
    xs = [x[:, a * C:(a + 1) * C].astype(np.float64) for a in range(4)]
    ws = [kernel[:, c * q:(c + 1) * q].astype(np.float64) for c in range(4)]
    outs = []
    for b in range(4):
        acc = 0
        for a in range(4):
            t = xs[a] @ ws[IDX_DENSE[a][b]]
            acc = acc - t if NEG_DENSE[a][b] else acc + t
        outs.append(acc)
    y = np.concatenate(outs, axis=-1)
    if bias is not None:
        y = y + bias.astype(np.float64)
    return ACTIVATIONS[activation](y).astype(np.float32)


# ----------------------------------------------------------------------------------------------------------------------
# Gradients (what TF autodiff yields through slice / neg / concat / conv / matmul / bias_add / relu).
# ----------------------------------------------------------------------------------------------------------------------
def fold_full_grad(g_full, in_q, n_out, transposed):
    """fold dL/dW_full (..., 4 in_q, 4 n_out) into the stored layout (..., in_q, 4 n_out)."""
    idx, neg = (IDX_DENSE, NEG_DENSE) if transposed else (IDX_CONV, NEG_CONV)
    out = np.zeros(g_full.shape[:-2] + (in_q, 4 * n_out), dtype=g_full.dtype)
    for a in range(4):
        for b in range(4):
            blk = g_full[..., a * in_q:(a + 1) * in_q, b * n_out:(b + 1) * n_out]
            c = idx[a][b]
            if neg[a][b]:
                out[..., c * n_out:(c + 1) * n_out] -= blk
            else:
                out[..., c * n_out:(c + 1) * n_out] += blk
    return out


def _act_grad(activation, z, dy):
    if activation in (None, "linear"):
        return dy
    if activation == "relu":
        return dy * (z > 0)
    raise NotImplementedError("oracle backward supports linear and relu only")


def qdense_backward(x, kernel, bias, units, activation, dy):
    """Returns (dx, dkernel, dbias) in fp64 for y = act(x @ W_full + b)."""
    q = units // 4
    in_q = kernel.shape[0]
    w_full = expand_dense_kernel(kernel.astype(np.float64), q)
    z = x.astype(np.float64) @ w_full + (0 if bias is None else bias.astype(np.float64))
    dz = _act_grad(activation, z, dy.astype(np.float64))
    dx = dz @ w_full.T
    g_full = x.astype(np.float64).T @ dz
    return dx, fold_full_grad(g_full, in_q, q, transposed=True), dz.sum(axis=0)


def qconv_backward(x, kernel, bias, filters, strides, padding, data_format, dilation_rate, activation, dy):
    """Returns (dx, dkernel, dbias) in fp64 for the quaternion convolution (any rank, any stride / dilation)."""
    rank = kernel.ndim - 2
    strides, dilation = _tup(strides, rank), _tup(dilation_rate, rank)
    in_q = kernel.shape[-2]
    w_full = expand_conv_kernel(kernel.astype(np.float64), filters)
    xs = np.moveaxis(x, 1, -1) if data_format == "channels_first" else x
    xs = xs.astype(np.float64)
    ksz = w_full.shape[:rank]
    pads, outs = [(0, 0)], []
    for a in range(rank):
        lo, hi, o = pad_amounts(xs.shape[1 + a], ksz[a], strides[a], dilation[a], padding)
        pads.append((lo, hi))
        outs.append(o)
    pads.append((0, 0))
    xp = np.pad(xs, pads)
    z = np.zeros((xs.shape[0],) + tuple(outs) + (4 * filters,))
    sls = {}
    for tap in np.ndindex(*ksz):
        sl = [slice(None)]
        for a in range(rank):
            st = tap[a] * dilation[a]
            sl.append(slice(st, st + (outs[a] - 1) * strides[a] + 1, strides[a]))
        sl.append(slice(None))
        sls[tap] = tuple(sl)
        z += xp[sls[tap]] @ w_full[tap]
    if bias is not None:
        z = z + bias.astype(np.float64)
    dyl = np.moveaxis(dy, 1, -1) if data_format == "channels_first" else dy
    dz = _act_grad(activation, z, dyl.astype(np.float64))
    dxp = np.zeros_like(xp)
    g_full = np.zeros_like(w_full)
    red = tuple(range(rank + 1))
    for tap in np.ndindex(*ksz):
        dxp[sls[tap]] += dz @ w_full[tap].T
        g_full[tap] = np.tensordot(xp[sls[tap]], dz, axes=(red, red))
    crop = [slice(None)] + [slice(p[0], p[0] + xs.shape[1 + a]) for a, p in enumerate(pads[1:-1])] + [slice(None)]
    dx = dxp[tuple(crop)]
    if data_format == "channels_first":
        dx = np.moveaxis(dx, -1, 1)
    return dx, fold_full_grad(g_full, in_q, filters, transposed=False), dz.sum(axis=red)


# ----------------------------------------------------------------------------------------------------------------------
# Initialisers (complexnn/init.py).  Same NumPy RNG call order: three global np.random.uniform draws for the imaginary
# axis, then RandomState(seed=1337) rayleigh + uniform for modulus and phase.
# ----------------------------------------------------------------------------------------------------------------------
def _quaternion_init(kernel_shape, fan_in, fan_out, criterion, seed):
    if criterion == "glorot":
        s = 1.0 / np.sqrt(2 * (fan_in + fan_out))
    elif criterion == "he":
        s = 1.0 / np.sqrt(2 * fan_in)
    else:
        raise ValueError("Invalid criterion: " + str(criterion))
    n = int(np.prod(kernel_shape))
    v_i = np.random.uniform(0.0, 1.0, n)
    v_j = np.random.uniform(0.0, 1.0, n)
    v_k = np.random.uniform(0.0, 1.0, n)
    # the reference squares NumPy *scalars* (libm pow), which is not always bit-identical to the vectorised x*x
    sq = lambda v: np.fromiter((e ** 2 for e in v), dtype=np.float64, count=n)
    norm = np.sqrt(sq(v_i) + sq(v_j) + sq(v_k)) + 0.0001
    v_i, v_j, v_k = (v_i / norm).reshape(kernel_shape), (v_j / norm).reshape(kernel_shape), (v_k / norm).reshape(kernel_shape)
    rng = np.random.RandomState(1337 if seed is None else seed)
    modulus = rng.rayleigh(scale=s, size=kernel_shape)
    phase = rng.uniform(low=-np.pi, high=np.pi, size=kernel_shape)
    return np.concatenate([modulus * np.cos(phase), modulus * v_i * np.sin(phase), modulus * v_j * np.sin(phase),
                           modulus * v_k * np.sin(phase)], axis=-1)


def qconv_init(kernel_size, in_q, filters, criterion="he", seed=None):
    """complexnn/init.py:48-93: returns kernel_size + (in_q, 4F) float64 (F2 in SURVEY: 4x wider than declared)."""
    kernel_shape = tuple(kernel_size) + (int(in_q), int(filters))
    rfs = int(np.prod(kernel_size))
    return _quaternion_init(kernel_shape, in_q * rfs, filters * rfs, criterion, seed)


def qdense_init(in_q, q_units, criterion="he", seed=None):
    """complexnn/init.py:115-155: returns (in_q, 4 q_units) float64."""
    return _quaternion_init((int(in_q), int(q_units)), in_q, q_units, criterion, seed)


# ----------------------------------------------------------------------------------------------------------------------
# Timed CPU baselines ("the reference's CPU path"): fp32, expansion inside the timed region as in the reference graph.
# ----------------------------------------------------------------------------------------------------------------------
def qconv1d_forward_f32(x, kernel, bias, filters, padding="same", relu=True):
    """fp32 NumPy-literal QuaternionConv1D forward, stride 1, channels_last: expand -> im2col -> sgemm -> bias -> relu."""
    w = expand_conv_kernel(kernel, filters)                     # [k, 4in_q, 4F]
    k = w.shape[0]
    lo, hi, out = pad_amounts(x.shape[1], k, 1, 1, padding)
    xp = np.pad(x, ((0, 0), (lo, hi), (0, 0)))
    cols = np.concatenate([xp[:, t:t + out, :] for t in range(k)], axis=-1)       # [B, T, k*4in_q]
    y = cols.reshape(-1, cols.shape[-1]) @ w.reshape(-1, w.shape[-1])
    y = y.reshape(x.shape[0], out, -1)
    if bias is not None:
        y += bias
    if relu:
        np.maximum(y, 0, out=y)
    return y


def qconv2d_forward_f32(x, kernel, bias, filters, relu=True):
    """fp32 NumPy-literal QuaternionConv2D forward, stride 1, `same`, channels_first: expand -> im2col -> sgemm -> bias ->
    relu (complexnn/conv.py:327-343 with K.conv2d restated as shifted slices + one GEMM per sample)."""
    w = expand_conv_kernel(kernel, filters)                     # [kh, kw, 4in_q, 4F]
    kh, kw = w.shape[:2]
    B, C, H, W = x.shape
    lo_h, hi_h, _ = pad_amounts(H, kh, 1, 1, "same")
    lo_w, hi_w, _ = pad_amounts(W, kw, 1, 1, "same")
    xp = np.pad(x, ((0, 0), (0, 0), (lo_h, hi_h), (lo_w, hi_w)))
    wt = np.ascontiguousarray(w.reshape(kh * kw * C, -1).T)     # [4F, taps*4in_q]
    y = np.empty((B, wt.shape[0], H, W), np.float32)
    for b in range(B):
        cols = np.concatenate([xp[b, :, i:i + H, j:j + W].reshape(C, H * W) for i in range(kh) for j in range(kw)], axis=0)
        y[b] = (wt @ cols).reshape(-1, H, W)
    if bias is not None:
        y += bias[None, :, None, None]
    if relu:
        np.maximum(y, 0, out=y)
    return y


def qdense_forward_f32(x, kernel, bias, units, relu=True):
    y = x @ expand_dense_kernel(kernel, units // 4)
    if bias is not None:
        y += bias
    if relu:
        np.maximum(y, 0, out=y)
    return y


# The same three paths with the real convolution / matmul handed to torch's CPU kernels (oneDNN / MKL) -- the closest thing
# in this image to what TensorFlow's CPU backend would run under the reference; the expansion stays inside the timed call.
def qconv1d_forward_torch_cpu(x, kernel, bias, filters, padding="same", relu=True):
    import torch
    w = expand_conv_kernel(kernel, filters)                                  # [k, 4in_q, 4F]
    k = w.shape[0]
    lo, hi, _ = pad_amounts(x.shape[1], k, 1, 1, padding)
    xt = torch.from_numpy(x).permute(0, 2, 1)                                # NCL view of the channels_last memory
    if lo or hi:
        xt = torch.nn.functional.pad(xt, (lo, hi))
    y = torch.nn.functional.conv1d(xt, torch.from_numpy(np.ascontiguousarray(w.transpose(2, 1, 0))),
                                   torch.from_numpy(bias) if bias is not None else None)
    if relu:
        y = torch.relu_(y)
    return y.permute(0, 2, 1).contiguous().numpy()


def qconv2d_forward_torch_cpu(x, kernel, bias, filters, relu=True):
    import torch
    w = expand_conv_kernel(kernel, filters)                                  # [kh, kw, 4in_q, 4F]
    kh, kw = w.shape[:2]
    lo_h, hi_h, _ = pad_amounts(x.shape[2], kh, 1, 1, "same")
    lo_w, hi_w, _ = pad_amounts(x.shape[3], kw, 1, 1, "same")
    xt = torch.nn.functional.pad(torch.from_numpy(x), (lo_w, hi_w, lo_h, hi_h))
    y = torch.nn.functional.conv2d(xt, torch.from_numpy(np.ascontiguousarray(w.transpose(3, 2, 0, 1))),
                                   torch.from_numpy(bias) if bias is not None else None)
    if relu:
        y = torch.relu_(y)
    return y.numpy()


def qdense_forward_torch_cpu(x, kernel, bias, units, relu=True):
    import torch
    y = torch.from_numpy(x) @ torch.from_numpy(expand_dense_kernel(kernel, units // 4))
    if bias is not None:
        y += torch.from_numpy(bias)
    if relu:
        y = torch.relu_(y)
    return y.numpy()


def qmacs_conv(batch, out_spatial, kernel_size, in_q, filters):
    return int(batch) * int(np.prod(out_spatial)) * int(np.prod(kernel_size)) * int(in_q) * int(filters)


def qmacs_dense(rows, in_q, q_units):
    return int(rows) * int(in_q) * int(q_units)
