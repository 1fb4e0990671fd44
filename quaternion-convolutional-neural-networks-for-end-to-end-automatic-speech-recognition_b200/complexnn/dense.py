"""Quaternion fully-connected layer -- host-side mirror of reference complexnn/dense.py.

`units` is the number of REAL outputs (4 * quaternion units, dense.py:74-75); the stored kernel is (in_q, units) with
the components blocked on the last axis and is always produced by qdense_init whatever `kernel_initializer` says
(dense.py:98-109).  The Hamilton table is the transpose of the convolution's: y = conj(W) (x) x
(dense.py:139-143).  `call` is one launch of the fused kernel (qnn_dense_forward) instead of the reference's
slice / negate / concatenate / K.dot / K.bias_add / activation node chain (dense.py:131-162).
"""
import numpy as np

from . import _ops
from ._layer import (InputSpec, Layer, get_activation, get_initializer, passthrough, serialize_activation,
                     serialize_object)
from .init import qdense_init


class QuaternionDense(Layer):
    def __init__(self, units, activation=None, use_bias=True, init_criterion="he", kernel_initializer="quaternion",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None,
                 kernel_constraint=None, bias_constraint=None, seed=None, **kwargs):
        if "input_shape" not in kwargs and "input_dim" in kwargs:
            kwargs["input_shape"] = (kwargs.pop("input_dim"),)
        super(QuaternionDense, self).__init__(**kwargs)
        self.units = units
        self.q_units = units // 4
        self.activation = get_activation(activation)
        self.use_bias = use_bias
        self.init_criterion = init_criterion
        self.kernel_initializer = kernel_initializer
        self.bias_initializer = get_initializer(bias_initializer)
        self.kernel_regularizer = passthrough(kernel_regularizer)
        self.bias_regularizer = passthrough(bias_regularizer)
        self.activity_regularizer = passthrough(activity_regularizer)
        self.kernel_constraint = passthrough(kernel_constraint)
        self.bias_constraint = passthrough(bias_constraint)
        self.seed = np.random.randint(1, 10e6) if seed is None else seed
        self.input_spec = InputSpec(ndim=2)
        self.supports_masking = True

    def build(self, input_shape):
        assert len(input_shape) == 2
        assert input_shape[-1] % 2 == 0
        input_dim = input_shape[-1] // 4
        self.kernel_init = qdense_init((input_dim, self.q_units), self.init_criterion)
        self.kernel = self.add_weight(shape=(input_dim, self.units), initializer=self.kernel_init, name="r",
                                      regularizer=self.kernel_regularizer, constraint=self.kernel_constraint)
        if self.use_bias:
            self.bias = self.add_weight(shape=(self.units,), initializer="zeros", name="bias",
                                        regularizer=self.bias_regularizer, constraint=self.bias_constraint)
        else:
            self.bias = None
        self.input_spec = InputSpec(ndim=2, axes={-1: 4 * input_dim})
        self.built = True

    def call(self, inputs):
        fused = self.activation.fused
        out = _ops.dense_forward(inputs, self.kernel, self.bias, self.units,
                                 self.activation.name if fused else "linear", packed=self._packed_kernels())
        return out if fused else self.activation(out)

    def _packed_kernels(self):
        """Packed kernel images of this layer (see complexnn/_ops.py: PackedKernels), rebuilt only when the kernel changes."""
        cache = getattr(self, "_packed", None)
        if cache is None:
            cache = self._packed = _ops.PackedKernels()
        return cache

    def backward(self, inputs, outputs, grad_outputs, need_input_grad=True, grad_kernel_out=None, grad_bias_out=None):
        """Returns (grad_inputs | None, grad_kernel, grad_bias | None); device tensors only."""
        if not self.activation.fused or self.activation.name not in ("linear", "relu"):
            raise NotImplementedError("backward supports linear and relu activations")
        return _ops.dense_backward(inputs, outputs, grad_outputs, self.kernel, self.bias is not None, self.units,
                                   self.activation.name, need_input_grad, grad_kernel_out, grad_bias_out,
                                   packed=self._packed_kernels())

    def compute_output_shape(self, input_shape):
        assert input_shape and len(input_shape) == 2
        assert input_shape[-1]
        output_shape = list(input_shape)
        output_shape[-1] = self.units
        return tuple(output_shape)

    def get_config(self):
        if self.kernel_initializer == "quaternion":
            ki = "quaternion"   # the reference puts the live qdense_init object here (not serialisable)
        else:
            ki = serialize_object(self.kernel_initializer)
        config = {
            "units": self.units,
            "activation": serialize_activation(self.activation),
            "use_bias": self.use_bias,
            "init_criterion": self.init_criterion,
            "kernel_initializer": ki,
            "bias_initializer": serialize_object(self.bias_initializer),
            "kernel_regularizer": serialize_object(self.kernel_regularizer),
            "bias_regularizer": serialize_object(self.bias_regularizer),
            "activity_regularizer": serialize_object(self.activity_regularizer),
            "kernel_constraint": serialize_object(self.kernel_constraint),
            "bias_constraint": serialize_object(self.bias_constraint),
            "seed": self.seed,
        }
        base = super(QuaternionDense, self).get_config()
        return dict(list(base.items()) + list(config.items()))
