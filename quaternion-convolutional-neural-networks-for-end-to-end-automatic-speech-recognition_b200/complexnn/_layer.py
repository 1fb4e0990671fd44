"""Host-side layer protocol the quaternion layers plug into: the slice of the Keras 2 `Layer` contract the reference
relies on -- constructor kwargs `name` / `input_shape` / `trainable` / `dtype`, `add_weight`, lazy
`build` on first `__call__`, `get_weights` / `set_weights`, `get_config` / `from_config` -- with weights held as fp32
NumPy masters plus a lazily uploaded device mirror (a torch CUDA tensor used purely as a device-memory handle)."""
import re

import numpy as np

_FLOATX = "float32"
_uids = {}


def unique_name(prefix):
    _uids[prefix] = _uids.get(prefix, 0) + 1
    return "%s_%d" % (prefix, _uids[prefix])


def to_snake_case(name):
    """Keras' rule for default layer names: QuaternionConv1D -> quaternion_conv1d."""
    intermediate = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub("([a-z])([A-Z])", r"\1_\2", intermediate).lower()


def normalize_data_format(value):
    if value is None:
        value = "channels_last"
    v = str(value).lower()
    if v not in ("channels_first", "channels_last"):
        raise ValueError('The `data_format` argument must be one of "channels_first", "channels_last". Received: '
                         + str(value))
    return v


def normalize_tuple(value, n, name):
    if isinstance(value, (int, np.integer)):
        return (int(value),) * n
    try:
        t = tuple(int(v) for v in value)
    except (TypeError, ValueError):
        raise ValueError("The `%s` argument must be a tuple of %d integers. Received: %s" % (name, n, value))
    if len(t) != n:
        raise ValueError("The `%s` argument must be a tuple of %d integers. Received: %s" % (name, n, value))
    return t


def normalize_padding(value):
    p = str(value).lower()
    if p not in ("valid", "same", "causal"):
        raise ValueError('The `padding` argument must be one of "valid", "same" (or "causal" for Conv1D). Received: '
                         + str(value))
    return p


def conv_output_length(input_length, filter_size, padding, stride, dilation=1):
    if input_length is None:
        return None
    dilated = filter_size + (filter_size - 1) * (dilation - 1)
    if padding in ("same", "causal"):
        out = input_length
    elif padding == "valid":
        out = input_length - dilated + 1
    else:
        raise ValueError("Invalid padding: " + str(padding))
    return (out + stride - 1) // stride


class SymbolicTensor(object):
    """Placeholder produced by `Input(...)` and by calling a layer on another placeholder (Keras functional API):
    carries only a shape and the (layer, inputs) pair that produces it.  A `Model` replays these nodes eagerly."""
    _keras_symbolic = True

    def __init__(self, shape, node=None, dtype=None, name=None):
        self.shape = tuple(shape)
        self._keras_shape = self.shape
        self._node = node
        self.dtype = dtype or _FLOATX
        self.name = name

    def __repr__(self):
        return "<SymbolicTensor shape=%s>" % (self.shape,)


def is_symbolic(x):
    return getattr(x, "_keras_symbolic", False) or (
        isinstance(x, (list, tuple)) and len(x) > 0 and all(getattr(t, "_keras_symbolic", False) for t in x))


class InputSpec(object):
    def __init__(self, dtype=None, shape=None, ndim=None, max_ndim=None, min_ndim=None, axes=None):
        self.dtype, self.shape, self.ndim = dtype, shape, ndim
        self.max_ndim, self.min_ndim, self.axes = max_ndim, min_ndim, dict(axes or {})


# ------------------------------------------------------------------------------------------------------------------
# identifier <-> object helpers (keras.activations / initializers / regularizers / constraints `.get` / `.serialize`)
# ------------------------------------------------------------------------------------------------------------------
FUSED_ACTIVATIONS = ("linear", "relu", "tanh", "sigmoid", "hard_sigmoid", "softplus", "softsign", "elu", "selu",
                     "exponential")


class Activation(object):
    """A named activation.  Names in FUSED_ACTIVATIONS run inside the kernel epilogue; `softmax` and user callables
    run as a separate pass on the layer output."""

    def __init__(self, name, fn=None):
        self.__name__ = self.name = name
        self.fn = fn

    @property
    def fused(self):
        return self.fn is None and self.name in FUSED_ACTIVATIONS

    def __call__(self, x):
        if self.fn is not None:
            return self.fn(x)
        if self.name == "linear":
            return x
        if self.name == "softmax":
            if isinstance(x, np.ndarray):
                e = np.exp(x - x.max(axis=-1, keepdims=True))
                return e / e.sum(axis=-1, keepdims=True)
            import torch
            return torch.softmax(x, dim=-1)
        raise NotImplementedError("activation %r only exists as a fused kernel epilogue" % self.name)

    def __eq__(self, other):
        return isinstance(other, Activation) and other.name == self.name and other.fn is self.fn

    def __hash__(self):
        return hash((self.name, self.fn))


def get_activation(identifier):
    if identifier is None:
        return Activation("linear")
    if isinstance(identifier, Activation):
        return identifier
    if isinstance(identifier, str):
        if identifier in FUSED_ACTIVATIONS or identifier == "softmax":
            return Activation(identifier)
        raise ValueError("Could not interpret activation function identifier: " + identifier)
    if callable(identifier):
        return Activation(getattr(identifier, "__name__", "custom"), identifier)
    raise ValueError("Could not interpret activation function identifier: " + str(identifier))


def serialize_activation(act):
    return act.__name__


class Initializer(object):
    def __call__(self, shape, dtype=None):
        raise NotImplementedError

    def get_config(self):
        return {}

    @classmethod
    def from_config(cls, config):
        return cls(**config)


class Zeros(Initializer):
    def __call__(self, shape, dtype=None):
        return np.zeros(shape, dtype=dtype or _FLOATX)


class Ones(Initializer):
    def __call__(self, shape, dtype=None):
        return np.ones(shape, dtype=dtype or _FLOATX)


class Constant(Initializer):
    def __init__(self, value=0.0):
        self.value = value

    def __call__(self, shape, dtype=None):
        return np.full(shape, self.value, dtype=dtype or _FLOATX)

    def get_config(self):
        return {"value": self.value}


class RandomUniform(Initializer):
    def __init__(self, minval=-0.05, maxval=0.05, seed=None):
        self.minval, self.maxval, self.seed = minval, maxval, seed

    def __call__(self, shape, dtype=None):
        return np.random.RandomState(self.seed).uniform(self.minval, self.maxval, shape).astype(dtype or _FLOATX)

    def get_config(self):
        return {"minval": self.minval, "maxval": self.maxval, "seed": self.seed}


class RandomNormal(Initializer):
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev, self.seed = mean, stddev, seed

    def __call__(self, shape, dtype=None):
        return np.random.RandomState(self.seed).normal(self.mean, self.stddev, shape).astype(dtype or _FLOATX)

    def get_config(self):
        return {"mean": self.mean, "stddev": self.stddev, "seed": self.seed}


_INITIALIZERS = {"zeros": Zeros, "ones": Ones, "constant": Constant, "random_uniform": RandomUniform,
                 "uniform": RandomUniform, "random_normal": RandomNormal, "normal": RandomNormal}
_INITIALIZERS.update({c.__name__: c for c in (Zeros, Ones, Constant, RandomUniform, RandomNormal)})


def get_initializer(identifier):
    if identifier is None:
        return None
    if isinstance(identifier, dict):
        return _INITIALIZERS[identifier["class_name"]](**identifier.get("config", {}))
    if isinstance(identifier, str):
        if identifier not in _INITIALIZERS:
            raise ValueError("Unknown initializer: " + identifier)
        return _INITIALIZERS[identifier]()
    if isinstance(identifier, type):
        return identifier()
    if callable(identifier):
        return identifier
    raise ValueError("Could not interpret initializer identifier: " + str(identifier))


def serialize_object(obj):
    """JSON-safe {'class_name', 'config'} (or None) for initializers / regularizers / constraints."""
    if obj is None:
        return None
    if isinstance(obj, (str, dict)):
        return obj
    cfg = obj.get_config() if hasattr(obj, "get_config") else {}
    name = obj.__name__ if isinstance(obj, type) else obj.__class__.__name__
    return {"class_name": name, "config": cfg}


def passthrough(identifier):
    """regularizers.get / constraints.get: objects are kept as given (they only matter to a training loop)."""
    return identifier


# ------------------------------------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------------------------------------
class Variable(object):
    """fp32 weight: NumPy master copy + device mirror uploaded on first use and after every assignment."""

    def __init__(self, value, name=None):
        self.name = name
        self._host = np.ascontiguousarray(np.asarray(value), dtype=np.float32)
        self._dev = None
        self._dev_is_master = False
        self.version = 0          # bumped by every assignment / optimiser step: keys caches of derived device images

    @property
    def shape(self):
        return self._host.shape

    def numpy(self):
        if self._dev_is_master:
            self._host = self._dev.detach().cpu().numpy().copy()
            self._dev_is_master = False
        return self._host

    def assign(self, value):
        value = np.asarray(value, dtype=np.float32)
        if value.shape != self._host.shape:
            raise ValueError("Layer weight shape %s not compatible with provided weight shape %s"
                             % (self._host.shape, value.shape))
        self._host = np.ascontiguousarray(value)
        self._dev_is_master = False
        self.version += 1
        if self._dev is not None:
            import torch
            with torch.no_grad():      # the mirror may be an autograd leaf (`parameter()`): in-place copy outside the graph
                self._dev.copy_(torch.from_numpy(self._host))

    def device(self, device="cuda"):
        """Device mirror (torch tensor).  In-place updates by an optimiser must call `mark_device_updated`."""
        import torch
        if self._dev is None or str(self._dev.device) != str(torch.device(device if device != "cuda" else
                                                                          "cuda:%d" % torch.cuda.current_device())):
            # moving to another device: the old mirror may hold the newest values (optimiser steps) -> pull them first
            grad = self._dev is not None and self._dev.requires_grad
            self._dev = torch.from_numpy(self.numpy()).to(device)
            if grad:
                self._dev.requires_grad_(True)
        return self._dev

    def mark_device_updated(self):
        self._dev_is_master = True
        self.version += 1

    def parameter(self, device="cuda"):
        """The device mirror as an autograd leaf (for an optimiser); call `mark_device_updated` after stepping it."""
        t = self.device(device)
        if not t.requires_grad:
            t.requires_grad_(True)
        return t

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return "<Variable %s shape=%s>" % (self.name, self.shape)


class Layer(object):
    _ALLOWED_KWARGS = {"input_shape", "batch_input_shape", "batch_size", "dtype", "name", "trainable", "weights",
                       "input_dtype"}

    def __init__(self, **kwargs):
        for k in kwargs:
            if k not in self._ALLOWED_KWARGS:
                raise TypeError("Keyword argument not understood:", k)
        self.name = kwargs.get("name") or unique_name(to_snake_case(self.__class__.__name__))
        self.trainable = kwargs.get("trainable", True)
        self.dtype = kwargs.get("dtype") or _FLOATX
        self.built = False
        self.input_spec = None
        self.supports_masking = False
        self._weights = []
        if "batch_input_shape" in kwargs:
            self.batch_input_shape = tuple(kwargs["batch_input_shape"])
        elif "input_shape" in kwargs:
            self.batch_input_shape = (kwargs.get("batch_size"),) + tuple(kwargs["input_shape"])
        self._initial_weights = kwargs.get("weights")

    # both Keras forms: add_weight(name=..., shape=..., initializer=...) and the legacy add_weight(shape, initializer=...)
    def add_weight(self, *args, **kwargs):
        args = list(args)
        name = kwargs.pop("name", None)
        shape = kwargs.pop("shape", None)
        if args and isinstance(args[0], str):
            name = args.pop(0)
        if args and shape is None:
            shape = args.pop(0)
        initializer = get_initializer(kwargs.pop("initializer", None) or "zeros")
        kwargs.pop("regularizer", None)
        kwargs.pop("constraint", None)
        kwargs.pop("trainable", None)
        kwargs.pop("dtype", None)
        # like Keras (K.variable(initializer(shape))) the initializer's return value decides the real shape: F2 in SURVEY
        var = Variable(initializer(tuple(shape)), name="%s/%s" % (self.name, name))
        self._weights.append(var)
        return var

    def build(self, input_shape):
        self.built = True

    def call(self, inputs):
        return inputs

    def compute_output_shape(self, input_shape):
        return input_shape

    def assert_input_compatibility(self, inputs):
        spec = self.input_spec
        if spec is None:
            return
        shape = tuple(inputs.shape)
        if spec.ndim is not None and len(shape) != spec.ndim:
            raise ValueError("Input 0 is incompatible with layer %s: expected ndim=%d, found ndim=%d"
                             % (self.name, spec.ndim, len(shape)))
        for axis, value in spec.axes.items():
            if value is not None and shape[int(axis)] is not None and shape[int(axis)] != value:
                raise ValueError("Input 0 is incompatible with layer %s: expected axis %s of input shape to have value %s "
                                 "but got shape %s" % (self.name, axis, value, shape))

    def __call__(self, inputs, **kwargs):
        if is_symbolic(inputs):
            return self._symbolic_call(inputs)
        self.assert_input_compatibility(inputs)
        if not self.built:
            self.build((None,) + tuple(int(s) for s in inputs.shape[1:]))
            self.built = True
            if self._initial_weights is not None:
                self.set_weights(self._initial_weights)
                self._initial_weights = None
            self.assert_input_compatibility(inputs)
        return self.call(inputs, **kwargs)

    def _symbolic_call(self, inputs):
        """Functional-API call on placeholders: build from the static shape, infer the output shape, record the node."""
        many = isinstance(inputs, (list, tuple))
        shapes = [tuple(t.shape) for t in inputs] if many else tuple(inputs.shape)
        if not many:
            self.assert_input_compatibility(inputs)
        if not self.built:
            self.build(shapes)
            self.built = True
            if self._initial_weights is not None:
                self.set_weights(self._initial_weights)
                self._initial_weights = None
        out_shape = self.compute_output_shape(shapes)
        return SymbolicTensor(out_shape, node=(self, inputs))

    @property
    def weights(self):
        return list(self._weights)

    @property
    def trainable_weights(self):
        return list(self._weights) if self.trainable else []

    @property
    def non_trainable_weights(self):
        return [] if self.trainable else list(self._weights)

    def get_weights(self):
        return [w.numpy().copy() for w in self._weights]

    def set_weights(self, weights):
        if len(weights) != len(self._weights):
            raise ValueError('You called `set_weights(weights)` on layer "%s" with a weight list of length %d, but the '
                             "layer was expecting %d weights." % (self.name, len(weights), len(self._weights)))
        for var, value in zip(self._weights, weights):
            var.assign(value)

    def count_params(self):
        return int(sum(int(np.prod(w.shape)) for w in self._weights))

    def get_config(self):
        config = {"name": self.name, "trainable": self.trainable}
        if hasattr(self, "batch_input_shape"):
            config["batch_input_shape"] = self.batch_input_shape
        if self.dtype:
            config["dtype"] = self.dtype
        return config

    @classmethod
    def from_config(cls, config):
        return cls(**config)
