"""Component getters -- mirrors reference complexnn/utils.py.  They codify the component-BLOCKED channel layout
[r | i | j | k]: part c of a tensor is the c-th quarter of its channel axis.  As in the reference (utils.py:17-79) the
channel axis is axis 1 ("channels_first" is hard-wired there) except for rank-3 tensors, where it is the last axis.
Works on NumPy arrays and torch tensors alike (pure views, no kernel)."""
from ._layer import Layer


def _part(x, c):
    ndim = len(x.shape)
    if ndim != 3:
        n = x.shape[1] // 4
        return x[:, c * n:(c + 1) * n] if c < 3 else x[:, 3 * n:]
    n = x.shape[-1] // 4
    return x[..., c * n:(c + 1) * n] if c < 3 else x[..., 3 * n:]


def get_rpart_first(x):
    return _part(x, 0)


def get_ipart_first(x):
    return _part(x, 1)


def get_jpart_first(x):
    return _part(x, 2)


def get_kpart_first(x):
    return _part(x, 3)


def getpart_quaternion_output_shape_first(input_shape):
    shape = list(input_shape)
    axis = -1 if len(shape) == 3 else 1
    shape[axis] = shape[axis] // 4
    return tuple(shape)


class _GetPart(Layer):
    _component = 0

    def call(self, inputs):
        return _part(inputs, self._component)

    def compute_output_shape(self, input_shape):
        return getpart_quaternion_output_shape_first(input_shape)


class GetRFirst(_GetPart):
    _component = 0


class GetIFirst(_GetPart):
    _component = 1


class GetJFirst(_GetPart):
    _component = 2


class GetKFirst(_GetPart):
    _component = 3
