"""Minimal `keras` facade: just enough of Keras 2's functional API for the reference's
models/example_model.py and working_example.py to import and run unchanged in an image where neither Keras nor
TensorFlow can be installed.  Execution is eager on torch CUDA tensors; the quaternion layers are the B200 kernels of
this repository (wrapped in a torch.autograd.Function for training), the few stock layers around them (Dense, Conv1D,
AveragePooling1D, Flatten, Dropout, softmax head, Adam, categorical cross-entropy) are plain torch ops -- they are
callers of the hot path, not the hot path.  Put `<package>/keras_facade` and `<package>` on sys.path.
"""
from . import backend, activations, initializers, regularizers, constraints, optimizers, losses, utils  # noqa: F401
from . import layers, models, callbacks  # noqa: F401

__version__ = "2.2.4-b200-facade"
