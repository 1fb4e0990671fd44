"""Minimal NumPy-backed stand-in for the Keras 2.x API surface that /root/reference/complexnn and
models/example_model.py touch.  TEST INFRASTRUCTURE ONLY (lives under oracle/): it exists so that the reference's
own, unmodified Python (slicing, sign table, concatenation order, initialisers) can be executed in a container
where neither Keras nor TensorFlow is installable, in order to generate the golden vectors under tests/golden/.
Semantics restated from the Keras 2.2 / TensorFlow 1.x documentation; nothing here is used by the product path.
"""
from . import backend, activations, initializers, regularizers, constraints, utils, layers, models, optimizers  # noqa

__version__ = "2.2.4-numpy-shim"
