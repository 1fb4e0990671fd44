class Optimizer(object):
    def __init__(self, **kwargs):
        self.config = kwargs


class Adam(Optimizer):
    pass


class SGD(Optimizer):
    pass


class RMSprop(Optimizer):
    pass
