from . import Layer


class Recurrent(Layer):
    pass
