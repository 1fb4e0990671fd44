import numpy as np
from . import backend as K


class Initializer(object):
    def __call__(self, shape, dtype=None):
        raise NotImplementedError

    def get_config(self):
        return {}

    @classmethod
    def from_config(cls, config):
        return cls(**config)


class Zeros(Initializer):
    def __call__(self, shape, dtype=None):
        return np.zeros(shape, dtype=dtype or K.floatx())


class Ones(Initializer):
    def __call__(self, shape, dtype=None):
        return np.ones(shape, dtype=dtype or K.floatx())


class RandomUniform(Initializer):
    def __init__(self, minval=-0.05, maxval=0.05, seed=None):
        self.minval, self.maxval, self.seed = minval, maxval, seed

    def __call__(self, shape, dtype=None):
        return np.random.RandomState(self.seed).uniform(self.minval, self.maxval, shape).astype(dtype or K.floatx())

    def get_config(self):
        return dict(minval=self.minval, maxval=self.maxval, seed=self.seed)


class VarianceScaling(Initializer):
    def __init__(self, scale=1.0, mode="fan_in", distribution="normal", seed=None):
        self.scale, self.mode, self.distribution, self.seed = scale, mode, distribution, seed

    def __call__(self, shape, dtype=None):
        fan_in, fan_out = _compute_fans(shape)
        n = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2.0}[self.mode]
        rng = np.random if self.seed is None else np.random.RandomState(self.seed)
        if self.distribution == "normal":
            w = rng.normal(0.0, np.sqrt(self.scale / max(1.0, n)), shape)
        else:
            lim = np.sqrt(3.0 * self.scale / max(1.0, n))
            w = rng.uniform(-lim, lim, shape)
        return w.astype(dtype or K.floatx())

    def get_config(self):
        return dict(scale=self.scale, mode=self.mode, distribution=self.distribution, seed=self.seed)


class Orthogonal(Initializer):
    def __init__(self, gain=1.0, seed=None):
        self.gain, self.seed = gain, seed


def glorot_uniform(seed=None):
    return VarianceScaling(1.0, "fan_avg", "uniform", seed)


def he_normal(seed=None):
    return VarianceScaling(2.0, "fan_in", "normal", seed)


_ALL = dict(zeros=Zeros, ones=Ones, random_uniform=RandomUniform, uniform=RandomUniform,
            glorot_uniform=glorot_uniform, he_normal=he_normal)


def _compute_fans(shape, data_format="channels_last"):
    if len(shape) == 2:
        return shape[0], shape[1]
    if len(shape) in (3, 4, 5):
        rfs = np.prod(shape[:-2])
        return shape[-2] * rfs, shape[-1] * rfs
    n = np.sqrt(np.prod(shape))
    return n, n


def get(identifier):
    if identifier is None:
        return None
    if isinstance(identifier, str):
        return _ALL[identifier.lower()]()
    if isinstance(identifier, type):
        return identifier()
    if callable(identifier):
        return identifier
    raise ValueError("Could not interpret initializer identifier: " + str(identifier))


def serialize(initializer):
    if initializer is None:
        return None
    return {"class_name": initializer.__class__.__name__, "config": initializer.get_config()}
