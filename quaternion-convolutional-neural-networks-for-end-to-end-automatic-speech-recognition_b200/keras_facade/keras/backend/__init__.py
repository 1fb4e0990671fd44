"""keras.backend on torch tensors (eager).  Only what complexnn-style code and the example models call."""
import numpy as np
import torch

_FLOATX = "float32"


def floatx():
    return _FLOATX


def epsilon():
    return 1e-7


def image_data_format():
    return "channels_last"


def normalize_data_format(value):
    from complexnn._layer import normalize_data_format as n
    return n(value)


def _t(x):
    return x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))


def shape(x):
    return tuple(x.shape)


def int_shape(x):
    return tuple(x.shape)


def ndim(x):
    return len(x.shape)


def constant(value, dtype=None, shape=None, name=None):
    return torch.full(tuple(shape) if shape is not None else (), float(value), dtype=getattr(torch, dtype or _FLOATX))


def variable(value, dtype=None, name=None):
    return torch.as_tensor(np.asarray(value), dtype=getattr(torch, dtype or _FLOATX))


def sqrt(x):
    return torch.sqrt(_t(x)) if torch.is_tensor(x) else float(np.sqrt(x))


def concatenate(tensors, axis=-1):
    return torch.cat([_t(t) for t in tensors], dim=axis)


def reshape(x, shape):
    return _t(x).reshape(tuple(int(s) for s in shape))


def dot(x, y):
    return _t(x) @ _t(y)


def bias_add(x, bias, data_format=None):
    x, bias = _t(x), _t(bias)
    if normalize_data_format(data_format) == "channels_first" and x.dim() > 2:
        return x + bias.reshape((1, -1) + (1,) * (x.dim() - 2))
    return x + bias


def relu(x, alpha=0.0, max_value=None):
    y = torch.nn.functional.leaky_relu(_t(x), alpha) if alpha else torch.relu(_t(x))
    return y if max_value is None else torch.clamp(y, max=max_value)


def softmax(x, axis=-1):
    return torch.softmax(_t(x), dim=axis)


def _same_pads(n, k, s, d):
    out = -(-n // s)
    total = max((out - 1) * s + (k - 1) * d + 1 - n, 0)
    return total // 2, total - total // 2


def _conv(x, kernel, strides, padding, data_format, dilation_rate, rank):
    """tf.nn.convolution semantics on torch: cross-correlation, kernel = spatial + (in, out)."""
    import torch.nn.functional as F
    x, kernel = _t(x), _t(kernel)
    cf = normalize_data_format(data_format) == "channels_first"
    if not cf:
        x = x.movedim(-1, 1)
    w = kernel.movedim(-1, 0).movedim(-1, 1)
    pads = []
    for a in reversed(range(rank)):
        if padding == "same":
            pads += list(_same_pads(x.shape[2 + a], kernel.shape[a], strides[a], dilation_rate[a]))
        elif padding == "causal":
            pads += [dilation_rate[a] * (kernel.shape[a] - 1), 0]
        else:
            pads += [0, 0]
    x = F.pad(x, pads)
    y = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[rank](x, w, stride=tuple(strides), dilation=tuple(dilation_rate))
    return y if cf else y.movedim(1, -1)


def conv1d(x, kernel, strides=1, padding="valid", data_format=None, dilation_rate=1):
    return _conv(x, kernel, (strides,), padding, data_format, (dilation_rate,), 1)


def conv2d(x, kernel, strides=(1, 1), padding="valid", data_format=None, dilation_rate=(1, 1)):
    return _conv(x, kernel, tuple(strides), padding, data_format, tuple(dilation_rate), 2)


def conv3d(x, kernel, strides=(1, 1, 1), padding="valid", data_format=None, dilation_rate=(1, 1, 1)):
    return _conv(x, kernel, tuple(strides), padding, data_format, tuple(dilation_rate), 3)


def ctc_batch_cost(*args, **kwargs):
    raise NotImplementedError("CTC is outside the quaternion conv/dense path")


def function(inputs, outputs, updates=None, **kwargs):
    """K.function(inputs, outputs): a callable taking a list of arrays and returning a list of NumPy arrays
    (models/interspeech_model.py:184 builds its validation function this way).  Runs the recorded layer graph eagerly on
    the CUDA device, like Model.predict."""
    from ..models import Model
    models = [Model(list(inputs), o) for o in outputs]

    def run(feeds):
        import torch
        outs = []
        with torch.no_grad():
            for m in models:
                outs.append(m._forward([m._to_device(f) for f in feeds], False).cpu().numpy())
        return outs

    return run
