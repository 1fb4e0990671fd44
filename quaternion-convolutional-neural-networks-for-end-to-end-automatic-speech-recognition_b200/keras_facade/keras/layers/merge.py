from . import Layer


class _Merge(Layer):
    pass
