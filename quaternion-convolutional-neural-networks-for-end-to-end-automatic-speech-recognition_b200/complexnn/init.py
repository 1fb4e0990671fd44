"""Quaternion weight initialisers -- host-side NumPy, mirrors reference complexnn/init.py.

Each weight is a quaternion in polar form  |w| (cos(phi) + u sin(phi))  with |w| ~ Rayleigh(s), phi ~ U(-pi, pi) and
u a random unit pure-imaginary quaternion; s follows the He (1/sqrt(2 fan_in)) or Glorot (1/sqrt(2 (fan_in + fan_out)))
criterion (init.py:61-66, 122-127).  Two properties are part of the contract with the layers:

  * layout: the four components are concatenated on the LAST axis, so the stored conv kernel is
    kernel_size + (in_q, 4*filters) although the layer asks for kernel_size + (in_q, filters)  (init.py:91)
  * RNG order: three draws from the *global* NumPy RNG for the axis u (positive octant, normalised with +1e-4), then
    RandomState(seed or 1337) for modulus and phase (init.py:70-86, 131-147).  After np.random.seed(n) the
    arrays are bit-identical to the reference's (tests/test_layers_cpu.py checks this against tests/golden/init.npz).
"""
import numpy as np

from ._layer import Initializer


def _fans(shape):
    """keras.initializers._compute_fans for channels_last kernels (spatial..., in, out)."""
    if len(shape) == 2:
        return shape[0], shape[1]
    receptive = int(np.prod(shape[:-2]))
    return shape[-2] * receptive, shape[-1] * receptive


def _scale(fan_in, fan_out, criterion):
    if criterion == "glorot":
        return 1.0 / np.sqrt(2 * (fan_in + fan_out))
    if criterion == "he":
        return 1.0 / np.sqrt(2 * fan_in)
    raise ValueError("Invalid criterion: " + str(criterion))


def _polar_quaternion(shape, s, seed):
    n = int(np.prod(shape))
    ui, uj, uk = (np.random.uniform(0.0, 1.0, n) for _ in range(3))
    # squares go through NumPy scalars (libm pow) exactly as the reference's per-weight loop does (init.py:74-78):
    # the vectorised x*x differs from pow(x, 2) in the last bit for ~0.1 % of the draws
    def sq(v):
        return np.fromiter((e ** 2 for e in v), dtype=np.float64, count=n)
    norm = np.sqrt(sq(ui) + sq(uj) + sq(uk)) + 0.0001
    ui, uj, uk = (ui / norm).reshape(shape), (uj / norm).reshape(shape), (uk / norm).reshape(shape)
    rng = np.random.RandomState(seed)
    modulus = rng.rayleigh(scale=s, size=shape)
    phase = rng.uniform(low=-np.pi, high=np.pi, size=shape)
    sin = np.sin(phase)
    return np.concatenate([modulus * np.cos(phase), modulus * ui * sin, modulus * uj * sin, modulus * uk * sin], axis=-1)


class qconv_init(Initializer):
    def __init__(self, kernel_size, input_dim, weight_dim, nb_filters=None, criterion="he", seed=None):
        assert len(kernel_size) == weight_dim and weight_dim in {0, 1, 2, 3}
        self.nb_filters = nb_filters
        self.kernel_size = kernel_size
        self.input_dim = input_dim
        self.weight_dim = weight_dim
        self.criterion = criterion
        self.seed = 1337 if seed is None else seed

    def __call__(self, shape, dtype=None):
        # `shape` is ignored, as in the reference: the initialiser knows its own geometry
        if self.nb_filters is not None:
            kernel_shape = tuple(self.kernel_size) + (int(self.input_dim), self.nb_filters)
        else:
            kernel_shape = (int(self.input_dim), self.kernel_size[-1])
        fan_in, fan_out = _fans(tuple(self.kernel_size) + (self.input_dim, self.nb_filters))
        return _polar_quaternion(kernel_shape, _scale(fan_in, fan_out, self.criterion), self.seed)

    def get_config(self):
        return {"kernel_size": tuple(self.kernel_size), "input_dim": int(self.input_dim), "weight_dim": self.weight_dim,
                "nb_filters": self.nb_filters, "criterion": self.criterion, "seed": self.seed}


class qdense_init(Initializer):
    def __init__(self, shape, criterion="he", seed=None):
        self.shape = shape
        self.criterion = criterion
        self.seed = 1337 if seed is None else seed

    def __call__(self, shape, dtype=None):
        return _polar_quaternion(tuple(self.shape), _scale(self.shape[0], self.shape[1], self.criterion), self.seed)

    def get_config(self):
        return {"shape": tuple(int(s) for s in self.shape), "criterion": self.criterion, "seed": self.seed}


class sqrt_init(Initializer):
    def __call__(self, shape, dtype=None):
        return np.full(shape, 1.0 / np.sqrt(2.0), dtype=dtype or "float32")
