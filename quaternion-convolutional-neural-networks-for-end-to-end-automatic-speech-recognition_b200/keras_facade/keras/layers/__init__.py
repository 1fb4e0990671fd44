"""keras.layers for the facade: `Layer` / `InputSpec` are the very classes the quaternion layers derive from; the stock
layers below are eager torch ops (autograd-tracked, weights in the same `Variable` containers)."""
import numpy as np
import torch
import torch.nn.functional as F

from complexnn._layer import (InputSpec, Layer, SymbolicTensor, Variable, conv_output_length, get_activation,  # noqa: F401
                              get_initializer, normalize_data_format, normalize_padding, normalize_tuple)
from .. import backend as K


def Input(shape=None, batch_shape=None, name=None, dtype=None, **kwargs):
    if batch_shape is not None:
        shape = tuple(batch_shape[1:])
    return SymbolicTensor((None,) + tuple(shape), node=None, dtype=dtype, name=name)


def _apply_activation(act, x):
    if act.fn is not None:
        return act.fn(x)
    name = act.name
    if name == "linear":
        return x
    fn = {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "softplus": F.softplus,
          "softsign": F.softsign, "elu": F.elu, "selu": F.selu, "exponential": torch.exp,
          "hard_sigmoid": lambda v: torch.clamp(0.2 * v + 0.5, 0.0, 1.0),
          "softmax": lambda v: torch.softmax(v, dim=-1)}[name]
    return fn(x)


class Lambda(Layer):
    def __init__(self, function, output_shape=None, **kwargs):
        super(Lambda, self).__init__(**kwargs)
        self.function, self._output_shape = function, output_shape

    def call(self, inputs):
        return self.function(inputs)

    def compute_output_shape(self, input_shape):
        if self._output_shape is None:
            return input_shape
        if callable(self._output_shape):
            return tuple(self._output_shape(input_shape))
        return (input_shape[0],) + tuple(self._output_shape)


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super(Activation, self).__init__(**kwargs)
        self.activation = get_activation(activation)

    def call(self, inputs):
        return _apply_activation(self.activation, inputs)


class Dropout(Layer):
    def __init__(self, rate, noise_shape=None, seed=None, **kwargs):
        super(Dropout, self).__init__(**kwargs)
        self.rate = rate
        self.training = False

    def call(self, inputs):
        return F.dropout(inputs, self.rate, training=True) if self.training and self.rate > 0 else inputs


SpatialDropout1D = Dropout


class Flatten(Layer):
    def call(self, inputs):
        return inputs.reshape(inputs.shape[0], -1)

    def compute_output_shape(self, input_shape):
        return (input_shape[0], int(np.prod(input_shape[1:])))


class Reshape(Layer):
    def __init__(self, target_shape, **kwargs):
        super(Reshape, self).__init__(**kwargs)
        self.target_shape = tuple(target_shape)

    def call(self, inputs):
        return inputs.reshape((inputs.shape[0],) + self.target_shape)

    def compute_output_shape(self, input_shape):
        return (input_shape[0],) + self.target_shape


class Permute(Layer):
    def __init__(self, dims, **kwargs):
        super(Permute, self).__init__(**kwargs)
        self.dims = tuple(dims)

    def call(self, inputs):
        return inputs.permute((0,) + self.dims).contiguous()

    def compute_output_shape(self, input_shape):
        return (input_shape[0],) + tuple(input_shape[d] for d in self.dims)


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None,
                 kernel_constraint=None, bias_constraint=None, **kwargs):
        if "input_shape" not in kwargs and "input_dim" in kwargs:
            kwargs["input_shape"] = (kwargs.pop("input_dim"),)
        super(Dense, self).__init__(**kwargs)
        self.units, self.use_bias = units, use_bias
        self.activation = get_activation(activation)
        self.kernel_initializer = get_initializer(kernel_initializer)
        self.bias_initializer = get_initializer(bias_initializer)

    def build(self, input_shape):
        self.kernel = self.add_weight(shape=(input_shape[-1], self.units), initializer=self.kernel_initializer, name="kernel")
        self.bias = self.add_weight(shape=(self.units,), initializer=self.bias_initializer, name="bias") if self.use_bias else None
        self.built = True

    def call(self, inputs):
        y = inputs @ self.kernel.device(inputs.device)
        if self.use_bias:
            y = y + self.bias.device(inputs.device)
        return _apply_activation(self.activation, y)

    def compute_output_shape(self, input_shape):
        return tuple(input_shape[:-1]) + (self.units,)


class _StockConv(Layer):
    rank = 1

    def __init__(self, filters, kernel_size, strides=1, padding="valid", data_format=None, dilation_rate=1,
                 activation=None, use_bias=True, kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None, kernel_constraint=None,
                 bias_constraint=None, **kwargs):
        super(_StockConv, self).__init__(**kwargs)
        self.filters = filters
        self.kernel_size = normalize_tuple(kernel_size, self.rank, "kernel_size")
        self.strides = normalize_tuple(strides, self.rank, "strides")
        self.dilation_rate = normalize_tuple(dilation_rate, self.rank, "dilation_rate")
        self.padding = normalize_padding(padding)
        self.data_format = normalize_data_format(data_format)
        self.activation = get_activation(activation)
        self.use_bias = use_bias
        self.kernel_initializer = get_initializer(kernel_initializer)
        self.bias_initializer = get_initializer(bias_initializer)

    def build(self, input_shape):
        cin = input_shape[1 if self.data_format == "channels_first" else -1]
        self.kernel = self.add_weight(shape=self.kernel_size + (cin, self.filters), initializer=self.kernel_initializer,
                                      name="kernel")
        self.bias = self.add_weight(shape=(self.filters,), initializer=self.bias_initializer, name="bias") if self.use_bias else None
        self.built = True

    def call(self, inputs):
        y = K._conv(inputs, self.kernel.device(inputs.device), self.strides, self.padding, self.data_format,
                    self.dilation_rate, self.rank)
        if self.use_bias:
            y = K.bias_add(y, self.bias.device(inputs.device), self.data_format)
        return _apply_activation(self.activation, y)

    def compute_output_shape(self, input_shape):
        cf = self.data_format == "channels_first"
        space = input_shape[2:] if cf else input_shape[1:-1]
        new = tuple(conv_output_length(space[a], self.kernel_size[a], self.padding, self.strides[a], self.dilation_rate[a])
                    for a in range(self.rank))
        return ((input_shape[0], self.filters) + new) if cf else ((input_shape[0],) + new + (self.filters,))


class Conv1D(_StockConv):
    rank = 1


class Conv2D(_StockConv):
    rank = 2


Convolution1D, Convolution2D = Conv1D, Conv2D


class AveragePooling1D(Layer):
    """tf.nn.avg_pool semantics: with SAME padding the divisor counts only the in-range samples."""

    def __init__(self, pool_size=2, strides=None, padding="valid", **kwargs):
        super(AveragePooling1D, self).__init__(**kwargs)
        self.pool_size = pool_size if isinstance(pool_size, int) else pool_size[0]
        self.strides = self.pool_size if strides is None else (strides if isinstance(strides, int) else strides[0])
        self.padding = normalize_padding(padding)

    def _pads(self, n):
        if self.padding == "valid":
            return 0, 0
        return K._same_pads(n, self.pool_size, self.strides, 1)

    def call(self, inputs):
        lo, hi = self._pads(inputs.shape[1])
        x = F.pad(inputs.transpose(1, 2), (lo, hi))
        y = F.avg_pool1d(x, self.pool_size, self.strides) * self.pool_size
        ones = F.pad(torch.ones((1, 1, inputs.shape[1]), device=inputs.device, dtype=inputs.dtype), (lo, hi))
        cnt = F.avg_pool1d(ones, self.pool_size, self.strides) * self.pool_size
        return (y / cnt).transpose(1, 2).contiguous()

    def compute_output_shape(self, input_shape):
        return (input_shape[0], conv_output_length(input_shape[1], self.pool_size, self.padding, self.strides), input_shape[2])


class TimeDistributed(Layer):
    """Applies `layer` to every time step: (B, T, ...) is folded to (B*T, ...) -- what the TIMIT model does around
    QuaternionDense (models/interspeech_model.py:149-157)."""

    def __init__(self, layer, **kwargs):
        super(TimeDistributed, self).__init__(**kwargs)
        self.layer = layer

    def build(self, input_shape):
        if not self.layer.built:
            self.layer.build((input_shape[0],) + tuple(input_shape[2:]))
            self.layer.built = True
        self._weights = self.layer._weights
        self.built = True

    def call(self, inputs):
        b, t = inputs.shape[0], inputs.shape[1]
        y = self.layer.call(inputs.reshape((b * t,) + tuple(inputs.shape[2:])).contiguous())
        return y.reshape((b, t) + tuple(y.shape[1:]))

    def compute_output_shape(self, input_shape):
        inner = self.layer.compute_output_shape((input_shape[0],) + tuple(input_shape[2:]))
        return (input_shape[0], input_shape[1]) + tuple(inner[1:])


class _Unsupported(Layer):
    def __init__(self, *a, **k):
        raise NotImplementedError(self.__class__.__name__ + " is outside the quaternion conv/dense path")


class MaxPooling2D(Layer):
    """tf.nn.max_pool semantics (SAME pads with -inf, odd element at the end)."""

    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", data_format=None, **kwargs):
        super(MaxPooling2D, self).__init__(**kwargs)
        self.pool_size = normalize_tuple(pool_size, 2, "pool_size")
        self.strides = normalize_tuple(self.pool_size if strides is None else strides, 2, "strides")
        self.padding = normalize_padding(padding)
        self.data_format = normalize_data_format(data_format)

    def call(self, inputs):
        x = inputs if self.data_format == "channels_first" else inputs.permute(0, 3, 1, 2)
        pads = []
        for a in (1, 0):
            pads += list(K._same_pads(x.shape[2 + a], self.pool_size[a], self.strides[a], 1)) if self.padding == "same" else [0, 0]
        y = F.max_pool2d(F.pad(x, pads, value=float("-inf")), self.pool_size, self.strides)
        return y if self.data_format == "channels_first" else y.permute(0, 2, 3, 1).contiguous()

    def compute_output_shape(self, input_shape):
        cf = self.data_format == "channels_first"
        space = input_shape[2:] if cf else input_shape[1:3]
        new = tuple(conv_output_length(space[a], self.pool_size[a], self.padding, self.strides[a]) for a in range(2))
        return ((input_shape[0], input_shape[1]) + new) if cf else ((input_shape[0],) + new + (input_shape[3],))


class AveragePooling2D(_Unsupported): pass
class AveragePooling3D(_Unsupported): pass
class BatchNormalization(_Unsupported): pass
class ConvLSTM2D(_Unsupported): pass

class PReLU(Layer):
    """keras.layers.PReLU: f(x) = max(x, 0) + alpha * min(x, 0) with a learned alpha per feature, shared over
    `shared_axes` (1-based, batch excluded).  models/interspeech_model.py:95-99 puts `PReLU(shared_axes=[1, 0])` after
    the quaternion convolutions; like Keras, axis 0 lands on `param_shape[-1]` (Python's negative index), which is what
    makes the variable-length time axis shareable there."""

    def __init__(self, alpha_initializer="zeros", alpha_regularizer=None, alpha_constraint=None, shared_axes=None, **kwargs):
        super(PReLU, self).__init__(**kwargs)
        self.alpha_initializer = get_initializer(alpha_initializer)
        if shared_axes is None:
            self.shared_axes = None
        elif not isinstance(shared_axes, (list, tuple)):
            self.shared_axes = [shared_axes]
        else:
            self.shared_axes = list(shared_axes)

    def build(self, input_shape):
        param_shape = list(input_shape[1:])
        for i in self.shared_axes or []:
            param_shape[i - 1] = 1
        if any(d is None for d in param_shape):
            raise ValueError("PReLU needs every unshared axis to be defined, got input shape %s" % (tuple(input_shape),))
        self.alpha = self.add_weight(shape=tuple(param_shape), initializer=self.alpha_initializer, name="alpha")
        self.built = True

    def call(self, inputs):
        return torch.relu(inputs) - self.alpha.device(inputs.device) * torch.relu(-inputs)

    def compute_output_shape(self, input_shape):
        return tuple(input_shape)

class Add(_Unsupported): pass
class Concatenate(_Unsupported): pass


def add(inputs, **kwargs):
    out = inputs[0]
    for t in inputs[1:]:
        out = out + t
    return out


def multiply(inputs, **kwargs):
    out = inputs[0]
    for t in inputs[1:]:
        out = out * t
    return out


def concatenate(inputs, axis=-1, **kwargs):
    return K.concatenate(inputs, axis)
