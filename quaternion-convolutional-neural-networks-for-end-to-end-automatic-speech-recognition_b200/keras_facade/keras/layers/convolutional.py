from . import Layer, Conv1D, Conv2D  # noqa: F401


class _Conv(Layer):
    pass
