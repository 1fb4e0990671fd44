import numpy as np
from . import backend as K


class Model(object):
    def __init__(self, inputs, outputs, name=None):
        self.inputs, self.outputs, self.name = inputs, outputs, name
        self.layers = []
        self._collect(outputs)

    def _collect(self, t):
        node = getattr(t, "_node", None)
        if node is None:
            return
        layer, parents, _ = node
        for p in parents if isinstance(parents, (list, tuple)) else [parents]:
            self._collect(p)
        if layer not in self.layers:
            self.layers.append(layer)

    def _eval(self, t, feed):
        if t is self.inputs:
            return feed
        layer, parents, kwargs = t._node
        if isinstance(parents, (list, tuple)):
            args = [self._eval(p, feed) for p in parents]
        else:
            args = self._eval(parents, feed)
        return K._t(layer.call(args, **kwargs))

    def predict(self, x, batch_size=None):
        return np.asarray(self._eval(self.outputs, K._t(np.asarray(x))))

    def summary(self):
        for l in self.layers:
            print("%-28s params=%d" % (l.name, l.count_params()))

    def get_weights(self):
        return [w for l in self.layers for w in l.get_weights()]

    def compile(self, *a, **k):
        self._compiled = (a, k)

    def fit(self, *a, **k):
        raise NotImplementedError("training is outside the oracle shim")


def load_model(*a, **k):
    raise NotImplementedError


def save_model(*a, **k):
    raise NotImplementedError
