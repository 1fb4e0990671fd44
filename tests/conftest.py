import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class G(object):
        def __init__(self):
            self._files = {}

        def load(self, name):
            if name not in self._files:
                self._files[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
            return self._files[name]

    return G()


@pytest.fixture(scope="session")
def native_lib():
    """Builds (if stale) and loads libqnn_b200.so; compute entry points need a GPU, symbol checks do not."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("qnn_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from complexnn import _native
    return _native.lib()
