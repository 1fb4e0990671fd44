from . import Layer


class _Conv(Layer):
    pass
