"""keras.backend on NumPy (eager).  Conv semantics = tf.nn.convolution: cross-correlation, kernel laid out
spatial..., C_in, C_out; SAME pads total = max((ceil(n/s)-1)*s + (k-1)*d + 1 - n, 0) with the odd element at the end."""
import numpy as np

_FLOATX = "float32"
_IMAGE_DATA_FORMAT = "channels_last"


class Tensor(np.ndarray):
    """ndarray that accepts attribute assignment (the reference sets `_keras_shape` on tensors)."""

    def __array_finalize__(self, obj):
        pass


def _t(a):
    return np.asarray(a).view(Tensor)


def floatx():
    return _FLOATX


def epsilon():
    return 1e-7


def image_data_format():
    return _IMAGE_DATA_FORMAT


def normalize_data_format(value):
    if value is None:
        value = image_data_format()
    v = value.lower()
    if v not in ("channels_first", "channels_last"):
        raise ValueError("The `data_format` argument must be one of "
                         '"channels_first", "channels_last". Received: ' + str(value))
    return v


def shape(x):
    return tuple(np.shape(x))


def int_shape(x):
    return tuple(np.shape(x))


def ndim(x):
    return np.ndim(x)


def constant(value, dtype=None, shape=None, name=None):
    return _t(np.full(shape if shape is not None else (), value, dtype=dtype or _FLOATX))


def variable(value, dtype=None, name=None):
    return _t(np.array(value, dtype=dtype or _FLOATX))


def sqrt(x):
    return np.sqrt(x)


def concatenate(tensors, axis=-1):
    return _t(np.concatenate([np.asarray(t) for t in tensors], axis=axis))


def dot(x, y):
    return _t(np.asarray(x) @ np.asarray(y))


def bias_add(x, bias, data_format=None):
    data_format = normalize_data_format(data_format)
    x = np.asarray(x)
    bias = np.asarray(bias)
    if data_format == "channels_first" and x.ndim > 2:
        return _t(x + bias.reshape((1, -1) + (1,) * (x.ndim - 2)))
    return _t(x + bias)


def _pad_amounts(n, k, s, d, padding):
    eff = (k - 1) * d + 1
    if padding == "valid":
        return 0, 0, (n - eff) // s + 1 if n >= eff else 0
    if padding == "same":
        out = -(-n // s)
        total = max((out - 1) * s + eff - n, 0)
        return total // 2, total - total // 2, out
    if padding == "causal":
        return d * (k - 1), 0, -(-n // s)
    raise ValueError("Invalid padding: " + str(padding))


def _conv_nd(x, kernel, strides, padding, data_format, dilation_rate, rank):
    data_format = normalize_data_format(data_format)
    x = np.asarray(x)
    kernel = np.asarray(kernel)
    if data_format == "channels_first":
        x = np.moveaxis(x, 1, -1)
    ksz = kernel.shape[:rank]
    cin, cout = kernel.shape[rank], kernel.shape[rank + 1]
    assert x.shape[-1] == cin, (x.shape, kernel.shape)
    pads, outs = [(0, 0)], []
    for i in range(rank):
        lo, hi, o = _pad_amounts(x.shape[1 + i], ksz[i], strides[i], dilation_rate[i], padding)
        pads.append((lo, hi))
        outs.append(o)
    pads.append((0, 0))
    xp = np.pad(x, pads)
    y = np.zeros((x.shape[0],) + tuple(outs) + (cout,), dtype=np.result_type(x, kernel))
    for tap in np.ndindex(*ksz):
        sl = [slice(None)]
        for i in range(rank):
            st = tap[i] * dilation_rate[i]
            sl.append(slice(st, st + (outs[i] - 1) * strides[i] + 1, strides[i]))
        sl.append(slice(None))
        y += xp[tuple(sl)] @ kernel[tap]
    if data_format == "channels_first":
        y = np.moveaxis(y, -1, 1)
    return _t(y)


def conv1d(x, kernel, strides=1, padding="valid", data_format=None, dilation_rate=1):
    return _conv_nd(x, kernel, (strides,), padding, data_format, (dilation_rate,), 1)


def conv2d(x, kernel, strides=(1, 1), padding="valid", data_format=None, dilation_rate=(1, 1)):
    return _conv_nd(x, kernel, tuple(strides), padding, data_format, tuple(dilation_rate), 2)


def conv3d(x, kernel, strides=(1, 1, 1), padding="valid", data_format=None, dilation_rate=(1, 1, 1)):
    return _conv_nd(x, kernel, tuple(strides), padding, data_format, tuple(dilation_rate), 3)


def relu(x, alpha=0.0, max_value=None):
    x = np.asarray(x)
    y = np.where(x > 0, x, alpha * x)
    if max_value is not None:
        y = np.minimum(y, max_value)
    return _t(y)


def softmax(x, axis=-1):
    x = np.asarray(x)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return _t(e / e.sum(axis=axis, keepdims=True))


def reshape(x, shape):
    return _t(np.reshape(np.asarray(x), tuple(int(v) for v in shape)))


def function(inputs, outputs, updates=None, **kwargs):
    """K.function([x], [y]) (models/interspeech_model.py:184): evaluates the recorded graph of y for a fed x."""
    from ..models import Model
    models = [Model(inputs[0] if len(inputs) == 1 else inputs, o) for o in outputs]
    return lambda feeds: [m.predict(feeds[0]) for m in models]


def ctc_batch_cost(y_true, y_pred, input_length, label_length):
    """CTC is outside the quaternion conv/dense path: a placeholder of the right shape, so that the reference's model
    builder (which wires the loss into the graph, models/interspeech_model.py:37-39,178) can run to completion.  Nothing
    reads its value."""
    return _t(np.zeros((np.shape(y_pred)[0], 1), dtype=_FLOATX))
