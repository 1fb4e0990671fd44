import numpy as np
from . import backend as K


def linear(x):
    return x


def relu(x, alpha=0.0, max_value=None):
    return K.relu(x, alpha, max_value)


def tanh(x):
    return K._t(np.tanh(x))


def sigmoid(x):
    return K._t(1.0 / (1.0 + np.exp(-np.asarray(x))))


def hard_sigmoid(x):
    return K._t(np.clip(0.2 * np.asarray(x) + 0.5, 0.0, 1.0))


def softmax(x, axis=-1):
    return K.softmax(x, axis)


def softplus(x):
    return K._t(np.logaddexp(np.asarray(x), 0.0))


def softsign(x):
    x = np.asarray(x)
    return K._t(x / (1.0 + np.abs(x)))


def elu(x, alpha=1.0):
    x = np.asarray(x)
    return K._t(np.where(x > 0, x, alpha * (np.exp(np.minimum(x, 0)) - 1.0)))


def selu(x):
    return K._t(1.0507009873554804934193349852946 * np.asarray(elu(x, 1.6732632423543772848170429916717)))


def exponential(x):
    return K._t(np.exp(x))


_ALL = dict(linear=linear, relu=relu, tanh=tanh, sigmoid=sigmoid, hard_sigmoid=hard_sigmoid, softmax=softmax,
            softplus=softplus, softsign=softsign, elu=elu, selu=selu, exponential=exponential)


def get(identifier):
    if identifier is None:
        return linear
    if callable(identifier):
        return identifier
    if identifier in _ALL:
        return _ALL[identifier]
    raise ValueError("Could not interpret activation function identifier: " + str(identifier))


def serialize(activation):
    return activation.__name__
