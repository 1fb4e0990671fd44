"""keras.optimizers: hyper-parameter holders; Model.compile turns them into torch.optim objects (Keras defaults:
Adam epsilon 1e-7 applied as in Keras, i.e. outside the square root -- same form torch uses)."""


class Optimizer(object):
    def __init__(self, lr=0.01, **kwargs):
        self.lr = lr
        self.kwargs = kwargs

    def build(self, params):
        raise NotImplementedError


class Adam(Optimizer):
    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=None, decay=0.0, amsgrad=False, **kwargs):
        super(Adam, self).__init__(lr, **kwargs)
        self.beta_1, self.beta_2, self.epsilon, self.amsgrad = beta_1, beta_2, epsilon or 1e-7, amsgrad

    def build(self, params):
        import torch
        return torch.optim.Adam(params, lr=self.lr, betas=(self.beta_1, self.beta_2), eps=self.epsilon,
                                amsgrad=self.amsgrad)


class SGD(Optimizer):
    def __init__(self, lr=0.01, momentum=0.0, decay=0.0, nesterov=False, **kwargs):
        super(SGD, self).__init__(lr, **kwargs)
        self.momentum, self.nesterov = momentum, nesterov

    def build(self, params):
        import torch
        return torch.optim.SGD(params, lr=self.lr, momentum=self.momentum, nesterov=self.nesterov and self.momentum > 0)


class RMSprop(Optimizer):
    def __init__(self, lr=0.001, rho=0.9, epsilon=None, decay=0.0, **kwargs):
        super(RMSprop, self).__init__(lr, **kwargs)
        self.rho, self.epsilon = rho, epsilon or 1e-7

    def build(self, params):
        import torch
        return torch.optim.RMSprop(params, lr=self.lr, alpha=self.rho, eps=self.epsilon)


def get(identifier):
    if isinstance(identifier, Optimizer):
        return identifier
    return {"adam": Adam, "sgd": SGD, "rmsprop": RMSprop}[str(identifier).lower()]()
