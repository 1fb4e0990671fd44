"""The minimal `keras` facade.  CPU part: the reference's own model files import unchanged on top of it
and build the expected graphs (skipped where /root/reference is absent, i.e. on the GPU box).  GPU part: the same
architectures run on the B200 kernels, reproduce the reference's outputs, and train."""
import os
import sys

import numpy as np
import pytest

from conftest import PKG

FACADE = os.path.join(PKG, "keras_facade")
REF = "/root/reference"


@pytest.fixture()
def facade_path():
    added = [p for p in (FACADE,) if p not in sys.path]
    for p in added:
        sys.path.insert(0, p)
    for m in [m for m in sys.modules if m == "keras" or m.startswith("keras.")]:
        if "keras_shim" in (getattr(sys.modules[m], "__file__", "") or ""):
            del sys.modules[m]
    yield
    for p in added:
        sys.path.remove(p)


def _qdnn(keras):
    """models/example_model.py:69-79, re-typed (the file itself cannot travel to the GPU box)."""
    from keras.layers import Input, Flatten, Dense, Dropout
    from keras.models import Model
    from complexnn import QuaternionDense
    inp = Input((250, 4))
    flat = Flatten()(inp)
    h0 = QuaternionDense(512, activation="relu")(flat)
    Dropout(0.3)(h0)
    h1 = QuaternionDense(512, activation="relu")(h0)
    Dropout(0.3)(h1)
    h2 = QuaternionDense(512, activation="relu")(h1)
    return Model(inp, Dense(8, activation="softmax")(h2))


def _qcnn(keras):
    """models/example_model.py:22-47, re-typed."""
    from keras.layers import Input, Flatten, Dense, AveragePooling1D
    from keras.models import Model
    from complexnn import QuaternionConv1D, QuaternionDense
    inp = Input((250, 4))
    c = QuaternionConv1D(32, 3, strides=1, activation="relu", padding="same")(inp)
    p = AveragePooling1D(2, padding="same")(c)
    c2 = QuaternionConv1D(64, 3, strides=1, activation="relu", padding="same")(p)
    p2 = AveragePooling1D(4, padding="same")(c2)
    d = QuaternionDense(256, activation="relu")(Flatten()(p2))
    return Model(inp, Dense(8, activation="softmax")(d))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_reference_model_files_import_and_build_unchanged(facade_path):
    sys.path.append(REF)
    try:
        for m in ("models", "models.example_model", "models.interspeech_model"):
            sys.modules.pop(m, None)
        import keras
        assert "b200-facade" in keras.__version__
        from models.example_model import CNN, DNN

        class P(object):
            pass
        p = P()
        p.model = "QCNN"
        qcnn = CNN(p)
        assert [l.__class__.__name__ for l in qcnn.layers] == ["QuaternionConv1D", "AveragePooling1D", "QuaternionConv1D",
                                                               "AveragePooling1D", "Flatten", "QuaternionDense", "Dense"]
        assert qcnn.count_params() == 551944 and qcnn.outputs.shape == (None, 8)
        assert [tuple(w.shape) for w in qcnn.layers[2].weights] == [(3, 32, 256), (256,)]
        p.model = "QDNN"
        qdnn = DNN(p)
        assert qdnn.count_params() == 264712
        # h1 takes h0, not the dropout output (example_model.py:75): dropouts are not on the path to the output
        assert [l.__class__.__name__ for l in qdnn.layers] == ["Flatten", "QuaternionDense", "QuaternionDense",
                                                               "QuaternionDense", "Dense"]
        p.model = "CNN"
        assert CNN(p).count_params() == 533128
        import models.interspeech_model as im           # Python-2 source (xrange, n/2): importable ...
        assert hasattr(im, "getTimitModel2D")

        class D(object):
            num_layers, start_filter, act, aact, dropout, l2, model, quat_init = 4, 8, "relu", "prelu", 0.1, 1e-4, \
                "quaternion", "quaternion"
        with pytest.raises(NameError):                  # ... not callable under Python 3 as it stands,
            im.getTimitModel2D(D())
        import builtins                                 # ... but it is with a Python-2 `xrange` stand-in and nothing else:
        builtins.xrange = lambda *a: range(*[int(v) for v in a])
        try:
            timit, val_function = im.getTimitModel2D(D())
        finally:
            del builtins.xrange
        names = [l.__class__.__name__ for l in timit.layers]
        assert names.count("QuaternionConv2D") == 5 and names.count("PReLU") == 8 and names.count("TimeDistributed") == 4
        assert names[:3] == ["QuaternionConv2D", "PReLU", "MaxPooling2D"] and names[-1] == "Lambda"     # ... -> CTC lambda
        assert [tuple(w.shape) for w in timit.layers[0].weights] == [(3, 5, 1, 32), (32,)]     # in_q = 1, 8 filters (F2)
        assert [tuple(w.shape) for w in timit.layers[1].weights] == [(1, 41, 1)]               # PReLU(shared_axes=[1, 0])
        assert timit.count_params() == 138338 and callable(val_function)
    finally:
        sys.path.remove(REF)
        for m in ("models", "models.example_model", "models.interspeech_model"):
            sys.modules.pop(m, None)


def test_functional_api_shapes(facade_path):
    import keras
    m = _qcnn(keras)
    assert m.outputs.shape == (None, 8) and m.count_params() == 551944
    m2 = _qdnn(keras)
    assert m2.count_params() == 264712
    lines = []
    m2.summary(print_fn=lines.append)
    assert any("quaternion_dense" in l and "(None, 512)" in l for l in lines)
    from keras.layers import Input, TimeDistributed
    from complexnn import QuaternionDense
    t = TimeDistributed(QuaternionDense(256))(Input((50, 128)))       # interspeech_model.py:149
    assert t.shape == (None, 50, 256)


@pytest.mark.gpu
def test_example_models_on_gpu_match_reference_outputs(facade_path, golden, native_lib):
    import keras
    g = golden.load("decoda_models")
    for tag, build in (("QDNN", _qdnn), ("QCNN", _qcnn)):
        model = build(keras)
        model.set_weights([g["%s.w%d" % (tag, j)] for j in range(len(model.get_weights()))])
        probs = model.predict(g["x"], batch_size=8)
        err = np.linalg.norm(probs - g[tag + ".probs"]) / np.linalg.norm(g[tag + ".probs"])
        assert err < 1e-3, (tag, err)


@pytest.mark.gpu
def test_quaternion_op_gradients_match_torch_autograd(facade_path, native_lib, monkeypatch):
    """Backward kernels vs torch autograd through the explicit Hamilton expansion (plain fp32 PyTorch reference)."""
    import torch
    import torch.nn.functional as F
    from keras.models import _QuaternionOp
    from complexnn import QuaternionConv1D, QuaternionDense
    monkeypatch.setenv("QNN_MATH", "fp32")
    monkeypatch.setenv("QNN_ALGO", "general")
    torch.manual_seed(0)
    np.random.seed(0)
    # conv1d
    conv = QuaternionConv1D(8, 3, padding="same", activation="relu")
    conv.build((None, 40, 16))
    conv.built = True
    x = torch.randn(5, 40, 16, device="cuda", requires_grad=True)
    k, b = conv.kernel.parameter("cuda"), conv.bias.parameter("cuda")
    y = _QuaternionOp.apply(x, k, b, conv, "relu")
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    got = [x.grad.clone(), k.grad.clone(), b.grad.clone()]
    kr, xr, br = k.detach().clone().requires_grad_(), x.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    fr, fi, fj, fk = kr[..., 0:8], kr[..., 8:16], kr[..., 16:24], kr[..., 24:32]
    w = torch.cat([torch.cat([fr, -fi, -fj, -fk], -2), torch.cat([fi, fr, -fk, fj], -2),
                   torch.cat([fj, fk, fr, -fi], -2), torch.cat([fk, -fj, fi, fr], -2)], -1)      # conv.py:327-331
    yr = torch.relu(F.conv1d(xr.transpose(1, 2), w.permute(2, 1, 0), padding=1).transpose(1, 2) + br)
    (yr * gy).sum().backward()
    for a, r in zip(got, [xr.grad, kr.grad, br.grad]):
        assert float((a - r).abs().max() / r.abs().max()) < 1e-4
    # dense
    dense = QuaternionDense(32, activation="relu")
    dense.build((None, 24))
    dense.built = True
    xd = torch.randn(11, 24, device="cuda", requires_grad=True)
    kd, bd = dense.kernel.parameter("cuda"), dense.bias.parameter("cuda")
    yd = _QuaternionOp.apply(xd, kd, bd, dense, "relu")
    gyd = torch.randn_like(yd)
    (yd * gyd).sum().backward()
    kdr, xdr, bdr = kd.detach().clone().requires_grad_(), xd.detach().clone().requires_grad_(), bd.detach().clone().requires_grad_()
    r_, i_, j_, k_ = kdr[:, 0:8], kdr[:, 8:16], kdr[:, 16:24], kdr[:, 24:32]
    wd = torch.cat([torch.cat([r_, -i_, -j_, -k_], -1), torch.cat([i_, r_, -k_, j_], -1),
                    torch.cat([j_, k_, r_, -i_], -1), torch.cat([k_, -j_, i_, r_], -1)], 0)       # dense.py:139-143
    ydr = torch.relu(xdr @ wd + bdr)
    (ydr * gyd).sum().backward()
    for a, r in zip([xd.grad, kd.grad, bd.grad], [xdr.grad, kdr.grad, bdr.grad]):
        assert float((a - r).abs().max() / r.abs().max()) < 1e-4


@pytest.mark.gpu
def test_fit_reduces_loss_like_working_example(facade_path, native_lib):
    """working_example.py:127-136 in miniature: compile(Adam(5e-4), categorical_crossentropy), fit, evaluate."""
    import keras
    from keras.optimizers import Adam
    rng = np.random.default_rng(0)
    n = 96
    labels = rng.integers(0, 8, n)
    x = rng.uniform(0, 0.2, size=(n, 250, 4)).astype(np.float32)
    x[:, :, 0] = 0.0                                            # DECODA: the real part is always 0
    for i, l in enumerate(labels):                              # plant a class-dependent pattern
        x[i, 30 * l:30 * l + 30, 1:] += 0.5
    y = np.eye(8, dtype=np.float32)[labels]
    np.random.seed(1)
    model = _qdnn(keras)
    model.compile(optimizer=Adam(lr=0.0005), loss="categorical_crossentropy", metrics=["accuracy"])
    loss0, acc0 = model.evaluate(x, y)
    hist = model.fit(x, y, validation_data=(x[:16], y[:16]), epochs=6, batch_size=3, verbose=0)
    loss1, acc1 = model.evaluate(x, y)
    assert loss1 < 0.5 * loss0 and acc1 > 0.9 and len(hist.history["loss"]) == 6
    assert set(hist.history) == {"loss", "acc", "val_loss", "val_acc"}


def test_prelu_matches_keras_definition_on_cpu(facade_path):
    """interspeech_model.py:95-99: PReLU(shared_axes=[1, 0]) after a channels_first quaternion conv with a variable-length
    time axis -- alpha has shape (1, 41, 1); f(x) = max(x, 0) + alpha * min(x, 0)."""
    import torch
    from keras.layers import Input, PReLU
    layer = PReLU(shared_axes=[1, 0])
    out = layer(Input((128, 41, None)))
    assert out.shape == (None, 128, 41, None)
    assert [tuple(w.shape) for w in layer.weights] == [(1, 41, 1)]
    alpha = np.linspace(-0.5, 0.5, 41).astype(np.float32).reshape(1, 41, 1)
    layer.set_weights([alpha])
    x = torch.randn(2, 128, 41, 7)
    y = layer.call(x).numpy()
    xn = x.numpy()
    np.testing.assert_allclose(y, np.maximum(xn, 0) + alpha[None] * np.minimum(xn, 0), rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        PReLU()(Input((4, None)))          # an unshared undefined axis cannot carry a weight
