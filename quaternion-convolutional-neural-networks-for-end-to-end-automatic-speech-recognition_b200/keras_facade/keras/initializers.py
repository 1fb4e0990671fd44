import numpy as np
from complexnn._layer import (Initializer, Zeros, Ones, Constant, RandomUniform, RandomNormal, _INITIALIZERS,  # noqa: F401
                              get_initializer, serialize_object as serialize)


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
        return w.astype(dtype or "float32")

    def get_config(self):
        return dict(scale=self.scale, mode=self.mode, distribution=self.distribution, seed=self.seed)


class Orthogonal(Initializer):
    def __init__(self, gain=1.0, seed=None):
        self.gain, self.seed = gain, seed


def glorot_uniform(seed=None):
    return VarianceScaling(1.0, "fan_avg", "uniform", seed)


def he_normal(seed=None):
    return VarianceScaling(2.0, "fan_in", "normal", seed)


_INITIALIZERS.update({"glorot_uniform": glorot_uniform, "he_normal": he_normal, "VarianceScaling": VarianceScaling})


def _compute_fans(shape, data_format="channels_last"):
    if len(shape) == 2:
        return shape[0], shape[1]
    rfs = int(np.prod(shape[:-2]))
    return shape[-2] * rfs, shape[-1] * rfs


def get(identifier):
    return get_initializer(identifier)
