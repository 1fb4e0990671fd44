"""keras.models for the facade: a functional `Model` that replays the recorded layer graph eagerly on torch CUDA
tensors.  `fit` is a plain mini-batch loop (shuffled epochs, Keras-style progress lines); gradients of the stock layers
come from torch autograd, gradients of the quaternion layers from this repository's backward kernels through
`_QuaternionOp`."""
import sys
import time

import numpy as np
import torch

from complexnn._layer import SymbolicTensor
from complexnn.conv import QuaternionConv
from complexnn.dense import QuaternionDense
from complexnn import _ops
from . import losses as _losses
from . import optimizers as _optimizers
from .layers import Dropout, TimeDistributed, _apply_activation


class _QuaternionOp(torch.autograd.Function):
    """y = layer(x) with the fused B200 kernel; backward = qnn_conv_backward / qnn_dense_backward."""

    @staticmethod
    def forward(ctx, x, kernel, bias, layer, act_name):
        x = x.contiguous()
        if isinstance(layer, QuaternionDense):
            y = _ops.dense_forward(x, layer.kernel, layer.bias, layer.units, act_name, packed=layer._packed_kernels())
        else:
            y = _ops.conv_forward(x, layer.kernel, layer.bias, layer.filters, layer.kernel_size, layer.strides,
                                  layer.padding, layer.data_format, layer.dilation_rate, act_name,
                                  packed=layer._packed_kernels())
        ctx.layer, ctx.act_name = layer, act_name
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        layer, dy = ctx.layer, dy.contiguous()
        if isinstance(layer, QuaternionDense):
            dx, dk, db = _ops.dense_backward(x, y, dy, layer.kernel, layer.bias is not None, layer.units, ctx.act_name,
                                             need_dx=ctx.needs_input_grad[0], packed=layer._packed_kernels())
        else:
            dx, dk, db = _ops.conv_backward(x, y, dy, layer.kernel, layer.bias is not None, layer.filters,
                                            layer.kernel_size, layer.strides, layer.padding, layer.data_format,
                                            layer.dilation_rate, ctx.act_name, need_dx=ctx.needs_input_grad[0],
                                            packed=layer._packed_kernels())
        return dx, dk, db, None, None


def _run_layer(layer, x, training):
    if isinstance(layer, Dropout):
        layer.training = training
    if isinstance(layer, TimeDistributed) and isinstance(layer.layer, (QuaternionConv, QuaternionDense)):
        b, t = x.shape[0], x.shape[1]
        y = _run_layer(layer.layer, x.reshape((b * t,) + tuple(x.shape[2:])), training)
        return y.reshape((b, t) + tuple(y.shape[1:]))
    if isinstance(layer, (QuaternionConv, QuaternionDense)) and torch.is_grad_enabled():
        dev = x.device
        kernel = layer.kernel.parameter(dev)
        bias = layer.bias.parameter(dev) if layer.bias is not None else None
        fused = layer.activation.fused and layer.activation.name in ("linear", "relu")
        y = _QuaternionOp.apply(x, kernel, bias, layer, layer.activation.name if fused else "linear")
        return y if fused else _apply_activation(layer.activation, y)
    return layer.call(x)


class Model(object):
    def __init__(self, inputs, outputs, name=None):
        self.inputs = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = outputs
        self.name = name or "model"
        self.nodes = []          # (layer, input symbolic tensor(s), output symbolic tensor) in execution order
        self._visit(outputs, set())
        self.layers = []
        for layer, _, _ in self.nodes:
            if layer not in self.layers:
                self.layers.append(layer)
        self.optimizer = self.loss = None
        self.metrics_names = ["loss"]
        self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None

    def _visit(self, t, seen):
        if id(t) in seen or not isinstance(t, SymbolicTensor) or t._node is None:
            return
        seen.add(id(t))
        layer, ins = t._node
        for p in ins if isinstance(ins, (list, tuple)) else [ins]:
            self._visit(p, seen)
        self.nodes.append((layer, ins, t))

    # ------------------------------------------------------------------------------------------------ execution
    def _forward(self, feeds, training):
        vals = {id(s): v for s, v in zip(self.inputs, feeds)}
        for layer, ins, out in self.nodes:
            x = [vals[id(p)] for p in ins] if isinstance(ins, (list, tuple)) else vals[id(ins)]
            vals[id(out)] = _run_layer(layer, x, training)
        return vals[id(self.outputs)]

    def _to_device(self, a):
        if self.device is None:
            raise RuntimeError("the keras facade executes on a CUDA device (no CPU fallback for the quaternion layers)")
        return torch.as_tensor(np.asarray(a, dtype=np.float32)).to(self.device)

    def predict(self, x, batch_size=32, verbose=0):
        outs = []
        with torch.no_grad():
            for i in range(0, len(x), batch_size):
                outs.append(self._forward([self._to_device(x[i:i + batch_size])], False).cpu().numpy())
        return np.concatenate(outs, axis=0)

    # ------------------------------------------------------------------------------------------------ training
    def compile(self, optimizer="adam", loss=None, metrics=None, **kwargs):
        self.optimizer = _optimizers.get(optimizer)
        self.loss = _losses.get(loss)
        self.metrics = list(metrics or [])
        self.metrics_names = ["loss"] + (["acc"] if "accuracy" in self.metrics or "acc" in self.metrics else [])
        self._torch_opt = None

    def _variables(self):
        seen, out = set(), []
        for layer in self.layers:
            if not layer.trainable:
                continue
            for v in layer.weights:
                if id(v) not in seen:
                    seen.add(id(v))
                    out.append(v)
        return out

    def _metrics(self, y_true, y_pred):
        vals = [float(self.loss(y_true, y_pred))]
        if len(self.metrics_names) > 1:
            vals.append(float((y_pred.argmax(-1) == y_true.argmax(-1)).float().mean()))
        return vals

    def train_on_batch(self, x, y):
        variables = self._variables()
        if self._torch_opt is None:
            self._torch_opt = self.optimizer.build([v.parameter(self.device) for v in variables])
        xb, yb = self._to_device(x), self._to_device(y)
        self._torch_opt.zero_grad(set_to_none=True)
        pred = self._forward([xb], True)
        loss = self.loss(yb, pred)
        loss.backward()
        self._torch_opt.step()
        for v in variables:
            v.mark_device_updated()
        with torch.no_grad():
            return self._metrics(yb, pred.detach())

    def evaluate(self, x, y, batch_size=32, verbose=1):
        tot, n = np.zeros(len(self.metrics_names)), 0
        with torch.no_grad():
            for i in range(0, len(x), batch_size):
                xb, yb = self._to_device(x[i:i + batch_size]), self._to_device(y[i:i + batch_size])
                m = self._metrics(yb, self._forward([xb], False))
                tot += np.array(m) * len(xb)
                n += len(xb)
        res = list(tot / max(n, 1))
        return res if len(res) > 1 else res[0]

    def fit(self, x, y, batch_size=32, epochs=1, verbose=1, validation_data=None, shuffle=True, **kwargs):
        history = {k: [] for k in self.metrics_names}
        n = len(x)
        if verbose:
            print("Train on %d samples%s" % (n, ", validate on %d samples" % len(validation_data[0]) if validation_data else ""))
        for epoch in range(epochs):
            t0 = time.time()
            order = np.random.permutation(n) if shuffle else np.arange(n)
            tot = np.zeros(len(self.metrics_names))
            for i in range(0, n, batch_size):
                idx = order[i:i + batch_size]
                tot += np.array(self.train_on_batch(x[idx], y[idx])) * len(idx)
            logs = dict(zip(self.metrics_names, tot / n))
            if validation_data is not None:
                val = self.evaluate(validation_data[0], validation_data[1], batch_size=max(batch_size, 32), verbose=0)
                val = val if isinstance(val, list) else [val]
                logs.update({"val_" + k: v for k, v in zip(self.metrics_names, val)})
            for k, v in logs.items():
                history.setdefault(k, []).append(float(v))
            if verbose:
                print("Epoch %d/%d\n - %ds - %s" % (epoch + 1, epochs, time.time() - t0,
                                                    " - ".join("%s: %.4f" % kv for kv in logs.items())))
                sys.stdout.flush()

        class History(object):
            pass
        h = History()
        h.history = history
        self.history = h          # Keras keeps the last History on the model
        return h

    # ------------------------------------------------------------------------------------------------ bookkeeping
    def summary(self, print_fn=print):
        print_fn("_" * 65)
        print_fn("%-29s%-26s%-10s" % ("Layer (type)", "Output Shape", "Param #"))
        print_fn("=" * 65)
        for s in self.inputs:
            print_fn("%-29s%-26s%-10d" % ("input (InputLayer)", str(s.shape), 0))
        total = 0
        for layer, _, out in self.nodes:
            n = layer.count_params()
            total += n
            print_fn("%-29s%-26s%-10d" % (("%s (%s)" % (layer.name, layer.__class__.__name__))[:28], str(out.shape), n))
        print_fn("=" * 65)
        print_fn("Total params: %s" % format(total, ","))
        print_fn("_" * 65)

    def count_params(self):
        return int(sum(l.count_params() for l in self.layers))

    def get_weights(self):
        return [w for l in self.layers for w in l.get_weights()]

    def set_weights(self, weights):
        k = 0
        for l in self.layers:
            n = len(l.weights)
            l.set_weights(weights[k:k + n])
            k += n
        self._torch_opt = None


def load_model(*a, **k):
    raise NotImplementedError("HDF5 model files need h5py / Keras, which are not available here")


def save_model(*a, **k):
    raise NotImplementedError("HDF5 model files need h5py / Keras, which are not available here")
