"""Generates tests/golden/*.npz by running the REFERENCE's own complexnn package (imported unmodified from
/root/reference) on top of the NumPy Keras stand-in in oracle/keras_shim.  Run in the build container only
(`python oracle/make_golden.py`); /root/reference does not exist on the GPU box, the committed .npz files travel.

Every case stores fp32 inputs / weights and the reference result evaluated in float64 (shim floatx = float64, fed with
the fp32-rounded values), so the file pins the reference's slicing / sign / concatenation semantics independent of
summation order."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "keras_shim"))
sys.path.insert(0, REF)

import keras.backend as K  # noqa: E402  (the shim)

K._FLOATX = "float64"
import complexnn as ref  # noqa: E402  (the reference package)
from complexnn.init import qconv_init as ref_qconv_init, qdense_init as ref_qdense_init  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")

sys.path.insert(0, HERE)
from _cases import CONV_CASES, DENSE_CASES  # noqa: E402


def f32(a):
    return np.asarray(a, dtype=np.float32)


def run_layer(layer, x, rng):
    layer(K._t(np.zeros((1,) + x.shape[1:])))          # build (weights via the reference initialisers)
    ws = [f32(w) for w in layer.get_weights()]
    if len(ws) > 1:
        ws[-1] = f32(rng.normal(0, 0.1, ws[-1].shape))  # non-zero bias so the bias path is exercised
    layer.set_weights([w.astype(np.float64) for w in ws])
    y = np.asarray(layer.call(K._t(x.astype(np.float64))))
    return ws, y


def main():
    os.makedirs(OUT, exist_ok=True)
    cls = {1: ref.QuaternionConv1D, 2: ref.QuaternionConv2D, 3: ref.QuaternionConv3D}
    conv = {}
    for i, (name, rank, xs, filters, ksz, kw) in enumerate(CONV_CASES):
        np.random.seed(100 + i)
        rng = np.random.default_rng(100 + i)
        x = f32(rng.normal(0, 1, xs))
        layer = cls[rank](filters, ksz, **kw)
        ws, y = run_layer(layer, x, rng)
        conv[name + ".x"] = x
        conv[name + ".kernel"] = ws[0]
        if len(ws) > 1:
            conv[name + ".bias"] = ws[1]
        conv[name + ".y"] = f32(y)
        assert tuple(y.shape[1:]) == tuple(layer.compute_output_shape((None,) + xs[1:])[1:]), name
        print("conv ", name, xs, "->", y.shape, "kernel", ws[0].shape)
    np.savez(os.path.join(OUT, "conv_forward.npz"), **conv)

    dense = {}
    for i, (name, xs, units, kw) in enumerate(DENSE_CASES):
        np.random.seed(200 + i)
        rng = np.random.default_rng(200 + i)
        x = f32(rng.normal(0, 1, xs))
        layer = ref.QuaternionDense(units, **kw)
        ws, y = run_layer(layer, x, rng)
        dense[name + ".x"] = x
        dense[name + ".kernel"] = ws[0]
        if len(ws) > 1:
            dense[name + ".bias"] = ws[1]
        dense[name + ".y"] = f32(y)
        print("dense", name, xs, "->", y.shape, "kernel", ws[0].shape)
    np.savez(os.path.join(OUT, "dense_forward.npz"), **dense)

    # initialisers: exact arrays after np.random.seed
    init = {}
    np.random.seed(7)
    init["conv1d_he"] = ref_qconv_init(kernel_size=(3,), input_dim=5, weight_dim=1, nb_filters=6, criterion="he")((3, 5, 6))
    np.random.seed(8)
    init["conv2d_glorot"] = ref_qconv_init(kernel_size=(2, 3), input_dim=4, weight_dim=2, nb_filters=3,
                                           criterion="glorot")((2, 3, 4, 3))
    np.random.seed(9)
    init["dense_he"] = ref_qdense_init((6, 5), "he")((6, 20))
    np.random.seed(10)
    init["dense_glorot"] = ref_qdense_init((4, 7), "glorot")((4, 28))
    np.savez(os.path.join(OUT, "init.npz"), **init)

    # KATs: single quaternion, 1x1 kernel
    kat = {}
    c = ref.QuaternionConv1D(1, 1)
    c(K._t(np.zeros((1, 1, 4))))
    c.set_weights([np.array([[[1, 2, 3, 4.0]]]), np.zeros(4)])
    kat["conv_w1234_x5678"] = np.asarray(c.call(K._t(np.array([[[5, 6, 7, 8.0]]]))))
    d = ref.QuaternionDense(4)
    d(K._t(np.zeros((1, 4))))
    d.set_weights([np.array([[1, 2, 3, 4.0]]), np.zeros(4)])
    kat["dense_w1234_x5678"] = np.asarray(d.call(K._t(np.array([[5, 6, 7, 8.0]]))))
    np.savez(os.path.join(OUT, "kat.npz"), **kat)
    print("KAT conv", kat["conv_w1234_x5678"].ravel(), "dense", kat["dense_w1234_x5678"].ravel())

    # config 1: the reference's own QDNN / QCNN builders (models/example_model.py) on DECODA test documents
    from models.example_model import DNN, CNN  # noqa: E402
    sys.path.insert(0, REPO)
    from oracle.decoda import load_decoda  # noqa: E402

    x_test, y_test = load_decoda(os.path.join(REF, "decoda", "250_TEST_Q.data"))

    class P(object):
        pass

    dec = {"x": f32(x_test[:24]), "labels": f32(y_test[:24])}
    for tag, builder in (("QDNN", DNN), ("QCNN", CNN)):
        np.random.seed(300)
        p = P()
        p.model = tag
        model = builder(p)
        ws = [f32(w) for w in model.get_weights()]
        k = 0
        for l in model.layers:
            n = len(l.get_weights())
            l.set_weights([w.astype(np.float64) for w in ws[k:k + n]])
            k += n
        out = model.predict(dec["x"].astype(np.float64))
        for j, w in enumerate(ws):
            dec["%s.w%d" % (tag, j)] = w
        dec[tag + ".probs"] = f32(out)
        print(tag, "layers", [l.name for l in model.layers], "out", out.shape, "argmax", out.argmax(-1)[:8])
    np.savez(os.path.join(OUT, "decoda_models.npz"), **dec)

    # config 3 (secondary reading): the reference's own TIMIT builder, models/interspeech_model.py:getTimitModel2D
    # -- Python-2 source, run unchanged with a stand-in for `xrange` -- QuaternionConv2D (3,5) channels_first stack with
    # PReLU(shared_axes=[1,0]), MaxPooling2D((1,3)), Permute + reshape, 3 x TimeDistributed(QuaternionDense(256)),
    # TimeDistributed(Dense(62, softmax)); evaluated through its own validation function K.function([I], [pred]).
    import builtins
    builtins.xrange = lambda *a: range(*[int(v) for v in a])
    try:
        import models.interspeech_model as im
    finally:
        pass

    class D(object):
        num_layers, start_filter, act, aact, dropout, l2, model, quat_init = 4, 8, "relu", "prelu", 0.1, 1e-4, \
            "quaternion", "quaternion"
    np.random.seed(400)
    rng = np.random.default_rng(400)
    timit, val_function = im.getTimitModel2D(D())
    del builtins.xrange
    tm = {"x": f32(rng.normal(0, 1, (2, 4, 41, 12)))}
    k = 0
    for l in timit.layers:
        ws = [f32(w) for w in l.get_weights()]
        if l.__class__.__name__ == "PReLU":                       # zeros-initialised: give the slopes something to do
            ws = [f32(rng.uniform(-0.3, 0.3, w.shape)) for w in ws]
        elif len(ws) == 2:
            ws[1] = f32(rng.normal(0, 0.1, ws[1].shape))          # non-zero biases
        l.set_weights([w.astype(np.float64) for w in ws])
        for w in ws:
            tm["w%d" % k] = w
            k += 1
    tm["layers"] = np.array([l.__class__.__name__ for l in timit.layers])
    tm["pred"] = f32(val_function([tm["x"].astype(np.float64)])[0])
    print("TIMIT", len(timit.layers), "layers,", k, "weight arrays, pred", tm["pred"].shape, "row sums", tm["pred"].sum(-1).ravel()[:3])
    np.savez(os.path.join(OUT, "timit_model.npz"), **tm)
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden bytes:", total)


if __name__ == "__main__":
    main()
