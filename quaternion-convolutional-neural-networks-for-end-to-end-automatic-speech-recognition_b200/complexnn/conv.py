"""Quaternion convolution layers -- host-side mirror of reference complexnn/conv.py.

Same class names, constructor signatures and defaults (conv.py:93-120, 480-498, 615-633, 751-769), same stored
weights ([kernel, (10 unused gammas when normalize_weight), bias]; kernel = kernel_size + (in_q, 4*filters)),
same output-shape rule (conv.py:347-372) and the same get_config key set (conv.py:374-403 minus the subclass pops) --
made JSON-safe, the shipped one raises NameError.

What differs is `call`: the reference slices r,i,j,k, negates, concatenates a 4in_q x 4F real kernel and hands it to
K.conv{1,2,3}d, K.bias_add and the activation as separate graph nodes on every call (conv.py:294-343).  Here the whole
of that is ONE launch of the fused Hamilton kernel in libqnn_b200.so (qnn_conv_forward); nothing is expanded.
"""
import numpy as np

from . import _ops
from ._layer import (InputSpec, Layer, get_activation, get_initializer, normalize_data_format, normalize_padding,
                     normalize_tuple, passthrough, serialize_activation, serialize_object)
from .init import qconv_init, sqrt_init

_QUATERNION_INIT_NAMES = ("complex", "complex_independent", "glorot_complex", "he_complex", "quaternion",
                          "quaternion_independent")


def sanitizedInitGet(init):
    if init in ["sqrt_init"]:
        return sqrt_init
    if init in _QUATERNION_INIT_NAMES:
        return init
    return get_initializer(init)


def sanitizedInitSer(init):
    if init in [sqrt_init] or isinstance(init, sqrt_init):
        return "sqrt_init"
    if isinstance(init, str):
        return init
    if isinstance(init, qconv_init):
        return "quaternion"
    return serialize_object(init)


class QuaternionConv(Layer):
    """Abstract n-D quaternion convolution (rank 1, 2 or 3).  Channel axes are component-blocked [r|i|j|k]."""

    def __init__(self, rank, filters, kernel_size, strides=1, padding="valid", data_format="channels_last",
                 dilation_rate=1, activation=None, use_bias=True, normalize_weight=False,
                 kernel_initializer="quaternion", bias_initializer="zeros", gamma_diag_initializer=sqrt_init,
                 gamma_off_initializer="zeros", kernel_regularizer=None, bias_regularizer=None,
                 gamma_diag_regularizer=None, gamma_off_regularizer=None, activity_regularizer=None,
                 kernel_constraint=None, bias_constraint=None, gamma_diag_constraint=None, gamma_off_constraint=None,
                 init_criterion="he", seed=None, spectral_parametrization=False, epsilon=1e-7, **kwargs):
        super(QuaternionConv, self).__init__(**kwargs)
        self.rank = rank
        self.filters = filters
        self.kernel_size = normalize_tuple(kernel_size, rank, "kernel_size")
        self.strides = normalize_tuple(strides, rank, "strides")
        self.padding = normalize_padding(padding)
        self.data_format = normalize_data_format(data_format)
        self.dilation_rate = normalize_tuple(dilation_rate, rank, "dilation_rate")
        self.activation = get_activation(activation)
        self.use_bias = use_bias
        self.normalize_weight = normalize_weight      # accepted and stored, never used in call
        self.init_criterion = init_criterion
        self.spectral_parametrization = spectral_parametrization
        self.epsilon = epsilon
        self.kernel_initializer = sanitizedInitGet(kernel_initializer)
        self.bias_initializer = sanitizedInitGet(bias_initializer)
        self.gamma_diag_initializer = sanitizedInitGet(gamma_diag_initializer)
        self.gamma_off_initializer = sanitizedInitGet(gamma_off_initializer)
        self.kernel_regularizer = passthrough(kernel_regularizer)
        self.bias_regularizer = passthrough(bias_regularizer)
        self.gamma_diag_regularizer = passthrough(gamma_diag_regularizer)
        self.gamma_off_regularizer = passthrough(gamma_off_regularizer)
        self.activity_regularizer = passthrough(activity_regularizer)
        self.kernel_constraint = passthrough(kernel_constraint)
        self.bias_constraint = passthrough(bias_constraint)
        self.gamma_diag_constraint = passthrough(gamma_diag_constraint)
        self.gamma_off_constraint = passthrough(gamma_off_constraint)
        self.seed = np.random.randint(1, 10e6) if seed is None else seed
        self.input_spec = InputSpec(ndim=self.rank + 2)

    def build(self, input_shape):
        channel_axis = 1 if self.data_format == "channels_first" else -1
        if input_shape[channel_axis] is None:
            raise ValueError("The channel dimension of the inputs should be defined. Found `None`.")
        input_dim = input_shape[channel_axis] // 4
        self.kernel_shape = self.kernel_size + (input_dim, self.filters)
        # only the string 'quaternion' selects an initialiser here; anything else is a KeyError (conv.py:167)
        init_cls = {"quaternion": qconv_init}[self.kernel_initializer]
        kern_init = init_cls(kernel_size=self.kernel_size, input_dim=input_dim, weight_dim=self.rank,
                             nb_filters=self.filters, criterion=self.init_criterion)
        self.kernel = self.add_weight(self.kernel_shape, initializer=kern_init, name="kernel",
                                      regularizer=self.kernel_regularizer, constraint=self.kernel_constraint)
        gammas = ("rr", "ri", "rj", "rk", "ii", "ij", "ik", "jj", "jk", "kk")
        if self.normalize_weight:
            gamma_shape = (input_dim * self.filters,)
            # the reference's own table (conv.py:183-256): rr, ii, jj and -- as shipped -- jk take the *diag*
            # initialiser / regulariser / constraint, kk takes the *off* ones; reproduced so that weight lists and
            # checkpoints stay interchangeable (the gammas are never read in call)
            for g in gammas:
                diag = g in ("rr", "ii", "jj", "jk")
                setattr(self, "gamma_" + g, self.add_weight(
                    shape=gamma_shape, name="gamma_" + g,
                    initializer=self.gamma_diag_initializer if diag else self.gamma_off_initializer,
                    regularizer=self.gamma_diag_regularizer if diag else self.gamma_off_regularizer,
                    constraint=self.gamma_diag_constraint if diag else self.gamma_off_constraint))
        else:
            for g in gammas:
                setattr(self, "gamma_" + g, None)
        if self.use_bias:
            self.bias = self.add_weight((4 * self.filters,), initializer=self.bias_initializer, name="bias",
                                        regularizer=self.bias_regularizer, constraint=self.bias_constraint)
        else:
            self.bias = None
        self.input_spec = InputSpec(ndim=self.rank + 2, axes={channel_axis: input_dim * 4})
        self.built = True

    def call(self, inputs):
        fused = self.activation.fused
        out = _ops.conv_forward(inputs, self.kernel, self.bias, self.filters, self.kernel_size, self.strides,
                                self.padding, self.data_format, self.dilation_rate,
                                self.activation.name if fused else "linear", packed=self._packed_kernels())
        return out if fused else self.activation(out)

    def _packed_kernels(self):
        """Packed kernel images of this layer (K-major tf32 core matrices the tensor-core kernels consume), rebuilt only
        when the kernel `Variable` changes."""
        cache = getattr(self, "_packed", None)
        if cache is None:
            cache = self._packed = _ops.PackedKernels()
        return cache

    def backward(self, inputs, outputs, grad_outputs, need_input_grad=True, grad_kernel_out=None, grad_bias_out=None):
        """Gradients TF autodiff would derive from the reference graph; device tensors only.
        Returns (grad_inputs | None, grad_kernel, grad_bias | None)."""
        if not self.activation.fused or self.activation.name not in ("linear", "relu"):
            raise NotImplementedError("backward supports linear and relu activations")
        return _ops.conv_backward(inputs, outputs, grad_outputs, self.kernel, self.bias is not None, self.filters,
                                  self.kernel_size, self.strides, self.padding, self.data_format, self.dilation_rate,
                                  self.activation.name, need_input_grad, grad_kernel_out, grad_bias_out,
                                  packed=self._packed_kernels())

    def compute_output_shape(self, input_shape):
        return _ops.conv_out_shape(tuple(input_shape), self.filters, self.kernel_size, self.strides, self.padding,
                                   self.data_format, self.dilation_rate)

    def get_config(self):
        config = {
            "rank": self.rank,
            "filters": self.filters,
            "kernel_size": self.kernel_size,
            "strides": self.strides,
            "padding": self.padding,
            "data_format": self.data_format,
            "dilation_rate": self.dilation_rate,
            "activation": serialize_activation(self.activation),
            "use_bias": self.use_bias,
            "normalize_weight": self.normalize_weight,
            "kernel_initializer": sanitizedInitSer(self.kernel_initializer),
            "bias_initializer": sanitizedInitSer(self.bias_initializer),
            "gamma_diag_initializer": sanitizedInitSer(self.gamma_diag_initializer),
            "gamma_off_initializer": sanitizedInitSer(self.gamma_off_initializer),
            "kernel_regularizer": serialize_object(self.kernel_regularizer),
            "bias_regularizer": serialize_object(self.bias_regularizer),
            "gamma_diag_regularizer": serialize_object(self.gamma_diag_regularizer),
            "gamma_off_regularizer": serialize_object(self.gamma_off_regularizer),
            "activity_regularizer": serialize_object(self.activity_regularizer),
            "kernel_constraint": serialize_object(self.kernel_constraint),
            "bias_constraint": serialize_object(self.bias_constraint),
            "gamma_diag_constraint": serialize_object(self.gamma_diag_constraint),
            "gamma_off_constraint": serialize_object(self.gamma_off_constraint),
            "init_criterion": self.init_criterion,
            "spectral_parametrization": self.spectral_parametrization,
        }
        base = super(QuaternionConv, self).get_config()
        return dict(list(base.items()) + list(config.items()))


def _subclass_init(rank):
    # The 1D/2D/3D wrappers expose a narrower signature and do NOT forward `seed` (conv.py:499-518).
    def __init__(self, filters, kernel_size, strides=1 if rank == 1 else (1,) * rank, padding="valid",
                 data_format="channels_last" if rank == 1 else None, dilation_rate=1 if rank == 1 else (1,) * rank,
                 activation=None, use_bias=True, kernel_initializer="quaternion", bias_initializer="zeros",
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None, kernel_constraint=None,
                 bias_constraint=None, seed=None, init_criterion="he", spectral_parametrization=False, **kwargs):
        QuaternionConv.__init__(self, rank=rank, filters=filters, kernel_size=kernel_size, strides=strides,
                                padding=padding, data_format=data_format, dilation_rate=dilation_rate,
                                activation=activation, use_bias=use_bias, kernel_initializer=kernel_initializer,
                                bias_initializer=bias_initializer, kernel_regularizer=kernel_regularizer,
                                bias_regularizer=bias_regularizer, activity_regularizer=activity_regularizer,
                                kernel_constraint=kernel_constraint, bias_constraint=bias_constraint,
                                init_criterion=init_criterion, spectral_parametrization=spectral_parametrization,
                                **kwargs)
    return __init__


class QuaternionConv1D(QuaternionConv):
    """1-D quaternion convolution over (batch, steps, 4*in_q); `padding` may also be "causal"."""
    __init__ = _subclass_init(1)

    def get_config(self):
        config = super(QuaternionConv1D, self).get_config()
        config.pop("rank")
        config.pop("data_format")
        return config


class QuaternionConv2D(QuaternionConv):
    """2-D quaternion convolution over (batch, rows, cols, 4*in_q) or (batch, 4*in_q, rows, cols)."""
    __init__ = _subclass_init(2)

    def get_config(self):
        config = super(QuaternionConv2D, self).get_config()
        config.pop("rank")
        return config


class QuaternionConv3D(QuaternionConv):
    """3-D quaternion convolution over (batch, d1, d2, d3, 4*in_q) or (batch, 4*in_q, d1, d2, d3)."""
    __init__ = _subclass_init(3)

    def get_config(self):
        config = super(QuaternionConv3D, self).get_config()
        config.pop("rank")
        return config


QuaternionConvolution1D = QuaternionConv1D
QuaternionConvolution2D = QuaternionConv2D
QuaternionConvolution3D = QuaternionConv3D
