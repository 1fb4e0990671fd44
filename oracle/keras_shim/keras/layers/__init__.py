"""keras.layers on NumPy: just enough of the Layer protocol (add_weight/build/call/__call__) plus the handful of stock
layers models/example_model.py composes.  Eager: `Input` yields a 1-sample zero tensor, every layer call computes
immediately and remembers (layer, parents) on its output so that `Model.predict` can replay the graph on real data."""
import numpy as np
from .. import backend as K
from .. import activations, initializers, regularizers, constraints


class InputSpec(object):
    def __init__(self, dtype=None, shape=None, ndim=None, max_ndim=None, min_ndim=None, axes=None):
        self.dtype, self.shape, self.ndim = dtype, shape, ndim
        self.max_ndim, self.min_ndim, self.axes = max_ndim, min_ndim, axes or {}


_UID = {}


def _unique_name(prefix):
    _UID[prefix] = _UID.get(prefix, 0) + 1
    return "%s_%d" % (prefix, _UID[prefix])


def _snake(name):
    out = []
    for i, ch in enumerate(name):
        if ch.isupper() and i and not name[i - 1].isupper():
            out.append("_")
        out.append(ch.lower())
    return "".join(out)


class Layer(object):
    def __init__(self, **kwargs):
        allowed = {"input_shape", "batch_input_shape", "batch_size", "dtype", "name", "trainable", "weights", "input_dtype"}
        for k in kwargs:
            if k not in allowed:
                raise TypeError("Keyword argument not understood:", k)
        self.name = kwargs.get("name") or _unique_name(_snake(self.__class__.__name__))
        self.trainable = kwargs.get("trainable", True)
        self.built = False
        self.input_spec = None
        self.supports_masking = False
        self._weights = []
        self.dtype = kwargs.get("dtype") or K.floatx()

    # Keras 2.2 accepts the legacy positional form add_weight(shape, initializer=...) used by complexnn/conv.py:175
    def add_weight(self, *args, **kwargs):
        name = kwargs.pop("name", None)
        shape = kwargs.pop("shape", None)
        args = list(args)
        if args and isinstance(args[0], str):
            name = args.pop(0)
        if args and shape is None:
            shape = args.pop(0)
        initializer = kwargs.pop("initializer", None)
        dtype = kwargs.pop("dtype", None) or K.floatx()
        init = initializers.get(initializer)
        value = K.variable(np.asarray(init(tuple(shape))), dtype=dtype)   # K.variable(initializer(shape)): no shape check
        value.name = name
        self._weights.append(value)
        return value

    def build(self, input_shape):
        self.built = True

    def call(self, inputs, **kwargs):
        return inputs

    def compute_output_shape(self, input_shape):
        return input_shape

    def _check_spec(self, x):
        spec = self.input_spec
        if spec is None:
            return
        for s, t in zip(spec if isinstance(spec, (list, tuple)) else [spec], x if isinstance(x, (list, tuple)) else [x]):
            if s is None:
                continue
            if s.ndim is not None and np.ndim(t) != s.ndim:
                raise ValueError("Input 0 is incompatible with layer %s: expected ndim=%d, found ndim=%d"
                                 % (self.name, s.ndim, np.ndim(t)))
            for axis, value in s.axes.items():
                if value is not None and np.shape(t)[int(axis)] not in (value, None):
                    raise ValueError("Input 0 is incompatible with layer %s: expected axis %s of input shape to have "
                                     "value %s but got shape %s" % (self.name, axis, value, np.shape(t)))

    def __call__(self, inputs, **kwargs):
        self._check_spec(inputs)
        if not self.built:
            if isinstance(inputs, (list, tuple)):
                shp = [(None,) + tuple(np.shape(t)[1:]) for t in inputs]
            else:
                shp = (None,) + tuple(np.shape(inputs)[1:])
            self.build(shp)
            self.built = True
            self._check_spec(inputs)
        out = K._t(self.call(inputs, **kwargs))
        out._node = (self, inputs, kwargs)
        return out

    @property
    def weights(self):
        return list(self._weights)

    trainable_weights = weights

    def get_weights(self):
        return [np.array(w) for w in self._weights]

    def set_weights(self, weights):
        assert len(weights) == len(self._weights)
        for w, v in zip(self._weights, weights):
            assert w.shape == np.shape(v), (w.shape, np.shape(v))
            w[...] = v

    def count_params(self):
        return int(sum(w.size for w in self._weights))

    def get_config(self):
        return {"name": self.name, "trainable": self.trainable}

    @classmethod
    def from_config(cls, config):
        return cls(**config)


def Input(shape=None, batch_shape=None, name=None, dtype=None, **kwargs):
    if batch_shape is not None:
        shape = tuple(batch_shape[1:])
    t = K._t(np.zeros((1,) + tuple(1 if s is None else s for s in shape), dtype=dtype or K.floatx()))
    t._node = None
    t._is_input = True
    return t


class Lambda(Layer):
    def __init__(self, function, output_shape=None, **kwargs):
        super(Lambda, self).__init__(**kwargs)
        self.function = function

    def call(self, inputs, **kwargs):
        return self.function(inputs)


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super(Activation, self).__init__(**kwargs)
        self.activation = activations.get(activation)

    def call(self, inputs):
        return self.activation(inputs)


class Dropout(Layer):
    def __init__(self, rate, noise_shape=None, seed=None, **kwargs):
        super(Dropout, self).__init__(**kwargs)
        self.rate = rate

    def call(self, inputs, training=None):
        return inputs   # inference mode


SpatialDropout1D = Dropout


class Flatten(Layer):
    def call(self, inputs):
        x = np.asarray(inputs)
        return x.reshape(x.shape[0], -1)


class Reshape(Layer):
    def __init__(self, target_shape, **kwargs):
        super(Reshape, self).__init__(**kwargs)
        self.target_shape = tuple(target_shape)

    def call(self, inputs):
        x = np.asarray(inputs)
        return x.reshape((x.shape[0],) + self.target_shape)


class Permute(Layer):
    def __init__(self, dims, **kwargs):
        super(Permute, self).__init__(**kwargs)
        self.dims = tuple(dims)

    def call(self, inputs):
        return np.transpose(np.asarray(inputs), (0,) + self.dims)


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None,
                 kernel_constraint=None, bias_constraint=None, **kwargs):
        if "input_shape" not in kwargs and "input_dim" in kwargs:
            kwargs["input_shape"] = (kwargs.pop("input_dim"),)
        super(Dense, self).__init__(**kwargs)
        self.units, self.use_bias = units, use_bias
        self.activation = activations.get(activation)
        self.kernel_initializer = initializers.get(kernel_initializer)

    def build(self, input_shape):
        self.kernel = self.add_weight(shape=(input_shape[-1], self.units), initializer=self.kernel_initializer, name="kernel")
        self.bias = self.add_weight(shape=(self.units,), initializer="zeros", name="bias") if self.use_bias else None
        self.built = True

    def call(self, inputs):
        y = K.dot(inputs, self.kernel)
        if self.use_bias:
            y = K.bias_add(y, self.bias)
        return self.activation(y)


class _ConvStock(Layer):
    rank = 1

    def __init__(self, filters, kernel_size, strides=1, padding="valid", data_format=None, dilation_rate=1,
                 activation=None, use_bias=True, kernel_initializer="glorot_uniform", **kwargs):
        kwargs = {k: v for k, v in kwargs.items() if k in ("name", "input_shape", "trainable")}
        super(_ConvStock, self).__init__(**kwargs)
        from ..utils import conv_utils
        self.filters = filters
        self.kernel_size = conv_utils.normalize_tuple(kernel_size, self.rank, "kernel_size")
        self.strides = conv_utils.normalize_tuple(strides, self.rank, "strides")
        self.dilation_rate = conv_utils.normalize_tuple(dilation_rate, self.rank, "dilation_rate")
        self.padding = conv_utils.normalize_padding(padding)
        self.data_format = K.normalize_data_format(data_format)
        self.activation = activations.get(activation)
        self.use_bias = use_bias
        self.kernel_initializer = initializers.get(kernel_initializer)

    def build(self, input_shape):
        cin = input_shape[1 if self.data_format == "channels_first" else -1]
        self.kernel = self.add_weight(shape=self.kernel_size + (cin, self.filters), initializer=self.kernel_initializer,
                                      name="kernel")
        self.bias = self.add_weight(shape=(self.filters,), initializer="zeros", name="bias") if self.use_bias else None
        self.built = True

    def call(self, inputs):
        y = K._conv_nd(inputs, self.kernel, self.strides, self.padding, self.data_format, self.dilation_rate, self.rank)
        if self.use_bias:
            y = K.bias_add(y, self.bias, self.data_format)
        return self.activation(y)


class Conv1D(_ConvStock):
    rank = 1


class Conv2D(_ConvStock):
    rank = 2


Convolution1D, Convolution2D = Conv1D, Conv2D


class AveragePooling1D(Layer):
    """tf.nn.avg_pool semantics: with SAME padding the divisor counts only the in-range samples."""

    def __init__(self, pool_size=2, strides=None, padding="valid", **kwargs):
        super(AveragePooling1D, self).__init__(**kwargs)
        self.pool_size = pool_size if isinstance(pool_size, int) else pool_size[0]
        self.strides = self.pool_size if strides is None else (strides if isinstance(strides, int) else strides[0])
        self.padding = padding.lower()

    def call(self, inputs):
        x = np.asarray(inputs)
        n, p, s = x.shape[1], self.pool_size, self.strides
        lo, hi, out = K._pad_amounts(n, p, s, 1, self.padding)
        y = np.zeros((x.shape[0], out, x.shape[2]), dtype=x.dtype)
        for o in range(out):
            a, b = max(o * s - lo, 0), min(o * s - lo + p, n)
            y[:, o] = x[:, a:b].mean(axis=1)
        return y


class MaxPooling2D(Layer):
    """tf.nn.max_pool semantics (SAME pads with -inf, the odd element at the end)."""

    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", data_format=None, **kwargs):
        super(MaxPooling2D, self).__init__(**kwargs)
        self.pool_size = (pool_size, pool_size) if isinstance(pool_size, int) else tuple(pool_size)
        strides = self.pool_size if strides is None else strides
        self.strides = (strides, strides) if isinstance(strides, int) else tuple(strides)
        self.padding = padding.lower()
        self.data_format = K.normalize_data_format(data_format)

    def call(self, inputs):
        x = np.asarray(inputs)
        if self.data_format == "channels_last":
            x = np.moveaxis(x, -1, 1)
        pads, outs = [], []
        for a in range(2):
            lo, hi, out = K._pad_amounts(x.shape[2 + a], self.pool_size[a], self.strides[a], 1, self.padding)
            pads.append((lo, hi))
            outs.append(out)
        xp = np.pad(x, ((0, 0), (0, 0)) + tuple(pads), constant_values=-np.inf)
        y = np.full(x.shape[:2] + tuple(outs), -np.inf, dtype=x.dtype)
        for i in range(self.pool_size[0]):
            for j in range(self.pool_size[1]):
                y = np.maximum(y, xp[:, :, i:i + (outs[0] - 1) * self.strides[0] + 1:self.strides[0],
                                     j:j + (outs[1] - 1) * self.strides[1] + 1:self.strides[1]])
        return y if self.data_format == "channels_first" else np.moveaxis(y, 1, -1)


class TimeDistributed(Layer):
    """Applies the wrapped layer to every time step: (B, T, ...) -> reshape (B*T, ...) -> layer -> (B, T, ...)."""

    def __init__(self, layer, **kwargs):
        super(TimeDistributed, self).__init__(**kwargs)
        self.layer = layer

    def call(self, inputs):
        x = np.asarray(inputs)
        y = np.asarray(self.layer(K._t(x.reshape((x.shape[0] * x.shape[1],) + x.shape[2:]))))
        return y.reshape((x.shape[0], x.shape[1]) + y.shape[1:])

    def get_weights(self):
        return self.layer.get_weights()

    def set_weights(self, weights):
        self.layer.set_weights(weights)

    def count_params(self):
        return self.layer.count_params()


class PReLU(Layer):
    """f(x) = max(x, 0) + alpha * min(x, 0); alpha per feature, shared over `shared_axes` (Keras' own index arithmetic:
    param_shape[i - 1] = 1, so axis 0 lands on the LAST axis)."""

    def __init__(self, alpha_initializer="zeros", alpha_regularizer=None, alpha_constraint=None, shared_axes=None, **kwargs):
        super(PReLU, self).__init__(**kwargs)
        self.alpha_initializer = initializers.get(alpha_initializer)
        self.shared_axes = None if shared_axes is None else list(shared_axes) if isinstance(shared_axes, (list, tuple)) \
            else [shared_axes]

    def build(self, input_shape):
        param_shape = list(input_shape[1:])
        for i in self.shared_axes or []:
            param_shape[i - 1] = 1
        self.alpha = self.add_weight(shape=tuple(param_shape), name="alpha", initializer=self.alpha_initializer)
        self.built = True

    def call(self, inputs):
        x = np.asarray(inputs)
        return np.maximum(x, 0) + np.asarray(self.alpha) * np.minimum(x, 0)


class _Stub(Layer):
    def __init__(self, *a, **k):
        raise NotImplementedError(self.__class__.__name__ + " is outside the quaternion conv/dense path")


class AveragePooling2D(_Stub): pass
class AveragePooling3D(_Stub): pass
class BatchNormalization(_Stub): pass
class ConvLSTM2D(_Stub): pass
class Add(_Stub): pass
class Concatenate(_Stub): pass


def add(inputs, **kwargs):
    return K._t(sum(np.asarray(t) for t in inputs))


def multiply(inputs, **kwargs):
    out = np.asarray(inputs[0])
    for t in inputs[1:]:
        out = out * np.asarray(t)
    return K._t(out)


def concatenate(inputs, axis=-1, **kwargs):
    return K.concatenate(inputs, axis)
