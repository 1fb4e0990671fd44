"""CPU tests of the oracle (oracle/qoracle.py): it must reproduce the golden vectors produced by the reference's own
code (oracle/make_golden.py), agree with an independent conv implementation (torch CPU, fp64) and with its own
direct 16-block formulation, and its gradients must pass finite differences."""
import os

import numpy as np
import pytest

from oracle import qoracle as O
from cases import CONV_CASES, DENSE_CASES, conv_kwargs


def rel_err(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))) / (np.max(np.abs(b)) + 1e-30))


def test_kat_hamilton_products(golden):
    kat = golden.load("kat")
    # (1+2i+3j+4k)(x)(5+6i+7j+8k) = -60+12i+30j+24k ; conj(w)(x)x = 70+0i-16j-8k   
    np.testing.assert_array_equal(kat["conv_w1234_x5678"].ravel(), [-60, 12, 30, 24])
    np.testing.assert_array_equal(kat["dense_w1234_x5678"].ravel(), [70, 0, -16, -8])
    w = np.array([[[1, 2, 3, 4.0]]], dtype=np.float32)
    x = np.array([[[5, 6, 7, 8.0]]], dtype=np.float32)
    np.testing.assert_array_equal(O.qconv_forward(x, w, None, 1).ravel(), [-60, 12, 30, 24])
    np.testing.assert_array_equal(O.qdense_forward(x[0], w[0], None, 4).ravel(), [70, 0, -16, -8])
    # identity weight leaves x unchanged; norm is multiplicative
    one = np.array([[[1, 0, 0, 0.0]]], dtype=np.float32)
    np.testing.assert_array_equal(O.qconv_forward(x, one, None, 1), x)
    y = O.qconv_forward(x, w, None, 1)
    assert abs(np.linalg.norm(y) - np.linalg.norm(w) * np.linalg.norm(x)) < 1e-4


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_oracle_matches_reference_golden(golden, case):
    name, rank, xs, filters, ksz, kw = case
    g = golden.load("conv_forward")
    k = conv_kwargs(rank, kw)
    x, kernel, bias = g[name + ".x"], g[name + ".kernel"], g.get(name + ".bias")
    assert (bias is not None) == k["use_bias"]
    y = O.qconv_forward(x, kernel, bias, filters, k["strides"], k["padding"], k["data_format"], k["dilation_rate"],
                        k["activation"])
    assert y.shape == g[name + ".y"].shape
    assert rel_err(y, g[name + ".y"]) < 2e-6
    ksz_t = (ksz,) * rank if isinstance(ksz, int) else tuple(ksz)
    assert y.shape == tuple(O.qconv_output_shape(x.shape, filters, ksz_t, k["strides"], k["padding"], k["data_format"],
                                                  k["dilation_rate"]))
    # independent formulation: 16 signed block convolutions, no expanded weight
    yd = O.qconv_forward_direct(x, kernel, bias, filters, k["strides"], k["padding"], k["data_format"],
                                k["dilation_rate"], k["activation"])
    assert rel_err(yd, g[name + ".y"]) < 2e-6


@pytest.mark.parametrize("case", DENSE_CASES, ids=[c[0] for c in DENSE_CASES])
def test_dense_oracle_matches_reference_golden(golden, case):
    name, xs, units, kw = case
    g = golden.load("dense_forward")
    x, kernel, bias = g[name + ".x"], g[name + ".kernel"], g.get(name + ".bias")
    y = O.qdense_forward(x, kernel, bias, units, kw.get("activation"))
    assert rel_err(y, g[name + ".y"]) < 2e-6
    yd = O.qdense_forward_direct(x, kernel, bias, units, kw.get("activation"))
    assert rel_err(yd, g[name + ".y"]) < 2e-6


@pytest.mark.parametrize("rank,padding,stride,dil,cf", [
    (1, "same", 1, 1, False), (1, "same", 2, 1, False), (1, "valid", 2, 2, True), (1, "causal", 1, 3, False),
    (2, "same", (2, 1), (1, 1), True), (2, "valid", (1, 1), (2, 1), False), (2, "same", (1, 1), (1, 2), False),
    (3, "same", (1, 2, 1), (1, 1, 1), False), (3, "valid", (1, 1, 1), (1, 2, 1), True)])
def test_real_conv_matches_torch(rank, padding, stride, dil, cf):
    """Third-party semantics (TF SAME/VALID/causal, no kernel flip) against torch's independent implementation."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(rank * 10 + len(padding))
    sp = {1: (13,), 2: (9, 8), 3: (5, 6, 4)}[rank]
    ksz = {1: (4,), 2: (3, 2), 3: (2, 3, 2)}[rank]
    cin, cout = 3, 5
    x = rng.normal(size=(2,) + sp + (cin,))
    w = rng.normal(size=ksz + (cin, cout))
    xin = np.moveaxis(x, -1, 1) if cf else x
    y = O.real_conv(xin, w, stride, padding, "channels_first" if cf else "channels_last", dil)
    y = np.moveaxis(y, 1, -1) if cf else y
    st, dl = O._tup(stride, rank), O._tup(dil, rank)
    pads = []
    for a in range(rank):
        lo, hi, _ = O.pad_amounts(sp[a], ksz[a], st[a], dl[a], padding)
        pads = [lo, hi] + pads          # F.pad wants the last axis first
    xt = F.pad(torch.from_numpy(np.moveaxis(x, -1, 1)), pads)
    wt = torch.from_numpy(np.moveaxis(np.moveaxis(w, -1, 0), -1, 1))     # (out, in, spatial...)
    yt = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[rank](xt, wt, stride=st, dilation=dl)
    np.testing.assert_allclose(y, np.moveaxis(yt.numpy(), 1, -1), rtol=1e-10, atol=1e-10)


def test_init_bit_exact_with_reference(golden):
    g = golden.load("init")
    np.random.seed(7)
    np.testing.assert_array_equal(O.qconv_init((3,), 5, 6, "he"), g["conv1d_he"])
    np.random.seed(8)
    np.testing.assert_array_equal(O.qconv_init((2, 3), 4, 3, "glorot"), g["conv2d_glorot"])
    np.random.seed(9)
    np.testing.assert_array_equal(O.qdense_init(6, 5, "he"), g["dense_he"])
    np.random.seed(10)
    np.testing.assert_array_equal(O.qdense_init(4, 7, "glorot"), g["dense_glorot"])
    assert g["conv1d_he"].shape == (3, 5, 24)      # F2: four times wider than the declared (3, 5, 6)
    with pytest.raises(ValueError):
        O.qconv_init((3,), 5, 6, "lecun")


def _fd_check(f, params, grads, rng, eps=1e-5, n=6):
    for p, g in zip(params, grads):
        for _ in range(n):
            idx = tuple(rng.integers(0, s) for s in p.shape)
            old = p[idx]
            p[idx] = old + eps
            up = f()
            p[idx] = old - eps
            dn = f()
            p[idx] = old
            assert abs((up - dn) / (2 * eps) - g[idx]) < 1e-5 * max(1.0, abs(g[idx])), (idx, (up - dn) / (2 * eps), g[idx])


@pytest.mark.parametrize("rank,padding,stride,dil,cf,act", [
    (1, "same", 1, 1, False, "relu"), (1, "causal", 2, 2, False, None), (2, "valid", (2, 1), (1, 2), True, "relu"),
    (2, "same", (1, 2), (1, 1), False, None), (3, "same", (1, 1, 2), (1, 1, 1), False, "relu")])
def test_conv_gradients_finite_difference(rank, padding, stride, dil, cf, act):
    rng = np.random.default_rng(5)
    sp = {1: (9,), 2: (6, 5), 3: (3, 4, 5)}[rank]
    ksz = {1: (3,), 2: (3, 2), 3: (2, 2, 3)}[rank]
    in_q, F = 2, 3
    x = rng.normal(size=(2,) + sp + (4 * in_q,))
    if cf:
        x = np.moveaxis(x, -1, 1).copy()
    kern = rng.normal(size=ksz + (in_q, 4 * F))
    bias = rng.normal(size=4 * F)
    fmt = "channels_first" if cf else "channels_last"
    y0 = O.qconv_forward(x, kern, bias, F, stride, padding, fmt, dil, act, out_dtype=None)
    gy = rng.normal(size=y0.shape)

    def loss():
        return float((O.qconv_forward(x, kern, bias, F, stride, padding, fmt, dil, act, out_dtype=None) * gy).sum())

    dx, dk, db = O.qconv_backward(x, kern, bias, F, stride, padding, fmt, dil, act, gy)
    _fd_check(loss, [x, kern, bias], [dx, dk, db], rng)


def test_dense_gradients_finite_difference():
    rng = np.random.default_rng(6)
    x, kern, bias = rng.normal(size=(5, 12)), rng.normal(size=(3, 8)), rng.normal(size=8)
    gy = rng.normal(size=(5, 8))

    def loss():
        return float((O.qdense_forward(x, kern, bias, 8, "relu", out_dtype=None) * gy).sum())

    dx, dk, db = O.qdense_backward(x, kern, bias, 8, "relu", gy)
    _fd_check(loss, [x, kern, bias], [dx, dk, db], rng)


def test_f32_baseline_paths_match_oracle():
    rng = np.random.default_rng(11)
    x = rng.normal(size=(3, 50, 32)).astype(np.float32)
    k = rng.normal(size=(3, 8, 64)).astype(np.float32)
    b = rng.normal(size=64).astype(np.float32)
    assert rel_err(O.qconv1d_forward_f32(x, k, b, 16), O.qconv_forward(x, k, b, 16, 1, "same", activation="relu")) < 1e-5
    xd = rng.normal(size=(20, 32)).astype(np.float32)
    kd = rng.normal(size=(8, 64)).astype(np.float32)
    assert rel_err(O.qdense_forward_f32(xd, kd, b, 64), O.qdense_forward(xd, kd, b, 64, "relu")) < 1e-5


def test_empty_and_ragged_shapes():
    k = np.ones((3, 1, 8), dtype=np.float32)
    assert O.qconv_forward(np.zeros((0, 5, 4), np.float32), k, None, 2).shape == (0, 3, 8)
    assert O.qconv_forward(np.zeros((2, 2, 4), np.float32), k, None, 2).shape == (2, 0, 8)     # shorter than the kernel
    assert O.qconv_forward(np.zeros((2, 2, 4), np.float32), k, None, 2, padding="same").shape == (2, 2, 8)
    assert O.conv_output_length(7, 3, "same", 2) == 4 and O.conv_output_length(7, 3, "valid", 2, 2) == 2


def test_timed_cpu_baselines_agree_with_the_oracle():
    """The fp32 restatements bench.py times as the CPU arm (NumPy im2col + sgemm, and torch-CPU / oneDNN conv on the
    expanded weight) compute the same thing as the fp64 oracle."""
    rng = np.random.default_rng(11)
    x = rng.normal(size=(3, 50, 16)).astype(np.float32)
    k = (rng.normal(size=(3, 4, 32)) * 0.2).astype(np.float32)
    b = rng.normal(0, 0.1, 32).astype(np.float32)
    for pad in ("same", "valid", "causal"):
        ref = O.qconv_forward(x, k, b, 8, 1, pad, "channels_last", 1, "relu")
        for fn in (O.qconv1d_forward_f32, O.qconv1d_forward_torch_cpu):
            np.testing.assert_allclose(fn(x, k, b, 8, pad, True), ref, rtol=1e-5, atol=1e-5)
    x2 = rng.normal(size=(2, 16, 9, 10)).astype(np.float32)
    k2 = (rng.normal(size=(3, 2, 4, 32)) * 0.2).astype(np.float32)
    ref2 = O.qconv_forward(x2, k2, b, 8, (1, 1), "same", "channels_first", (1, 1), "relu")
    for fn in (O.qconv2d_forward_f32, O.qconv2d_forward_torch_cpu):
        np.testing.assert_allclose(fn(x2, k2, b, 8, True), ref2, rtol=1e-5, atol=1e-5)
    xd = rng.normal(size=(7, 16)).astype(np.float32)
    kd = (rng.normal(size=(4, 32)) * 0.2).astype(np.float32)
    refd = O.qdense_forward(xd, kd, b, 32, "relu")
    for fn in (O.qdense_forward_f32, O.qdense_forward_torch_cpu):
        np.testing.assert_allclose(fn(xd, kd, b, 32, True), refd, rtol=1e-5, atol=1e-5)


def timit_oracle_forward(g, conv=None, dense=None):
    """models/interspeech_model.py:getTimitModel2D (quaternion variant, n = 4 layers, sf = 8, PReLU) restated layer by layer
    on the weights of tests/golden/timit_model.npz.  `conv` / `dense` default to the oracle; the GPU test passes the
    product layers' forward instead."""
    conv = conv or (lambda h, k, b, F: O.qconv_forward(h, k, b, F, (1, 1), "same", "channels_first", (1, 1), None))
    dense = dense or (lambda h, k, b: O.qdense_forward(h, k, b, 256, None))
    prelu = lambda h, a: np.maximum(h, 0) + a * np.minimum(h, 0)
    w = [g["w%d" % i] for i in range(26)]
    h = prelu(conv(g["x"], w[0], w[1], 8), w[2][None])
    # MaxPooling2D((1, 3), 'same') with Keras' DEFAULT data_format (channels_last) on a channels_first tensor: it pools
    # the 41-feature axis by 3 (SAME: one -inf column at the end), exactly what the reference model does
    hp = np.pad(h, ((0, 0), (0, 0), (0, 1), (0, 0)), constant_values=-np.inf)
    h = hp.reshape(h.shape[0], h.shape[1], 14, 3, h.shape[3]).max(axis=3)
    i = 3
    for F in (8, 8, 16, 16):
        h = prelu(conv(h, w[i], w[i + 1], F), w[i + 2][None])
        i += 3
    B, C, Fq, T = h.shape
    h = np.transpose(h, (0, 3, 1, 2)).reshape(B * T, C * Fq)
    for _ in range(3):
        h = prelu(dense(np.ascontiguousarray(h, dtype=np.float32), w[i], w[i + 1]), w[i + 2])
        i += 3
    z = h.astype(np.float64) @ w[i] + w[i + 1]
    e = np.exp(z - z.max(-1, keepdims=True))
    return (e / e.sum(-1, keepdims=True)).reshape(B, T, 62)


def test_timit_model_oracle_matches_the_reference_builder(golden):
    """The oracle chained as getTimitModel2D chains its layers reproduces what the reference's own builder computed
    (through its validation function) on the Keras stand-in: pins QuaternionConv2D channels_first (3,5) 'same' with
    in_q = 1 / 8 / 16 and TimeDistributed(QuaternionDense) with in_q = 224 / 64."""
    g = golden.load("timit_model")
    assert list(g["layers"][:3]) == ["QuaternionConv2D", "PReLU", "MaxPooling2D"] and g["w0"].shape == (3, 5, 1, 32)
    pred = timit_oracle_forward(g)
    np.testing.assert_allclose(pred, g["pred"], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(pred.sum(-1), 1.0, rtol=1e-6)


# ----------------------------------------------------------------------------------------------------------------------
# The plain-C restatement (oracle/qoracle_c.c): a third, independent implementation of the reference path
# ----------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c_oracle():
    import ctypes
    import importlib.util
    from conftest import REPO
    spec = importlib.util.spec_from_file_location("graft_entry", os.path.join(REPO, "__graft_entry__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = ctypes.CDLL(mod.build_c_oracle())
    fp, i = ctypes.POINTER(ctypes.c_float), ctypes.c_int
    lib.qoc_conv1d_forward.argtypes = [fp, fp, fp, fp] + [i] * 9
    lib.qoc_conv2d_cf_forward.argtypes = [fp, fp, fp, fp] + [i] * 13
    lib.qoc_dense_forward.argtypes = [fp, fp, fp, fp] + [i] * 4
    lib.qoc_conv1d_out_len.argtypes = [i] * 5
    return lib


def _fp(a):
    import ctypes
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


PADS = {"valid": 0, "same": 1, "causal": 2}


def test_c_oracle_matches_reference_golden_and_numpy_oracle(c_oracle, golden):
    g = golden.load("conv_forward")
    checked = 0
    for name, rank, xs, filters, ksz, kw in CONV_CASES:
        k = conv_kwargs(rank, kw)
        if k["activation"] not in (None, "relu"):
            continue
        x, kern, bias = g[name + ".x"], g[name + ".kernel"], g.get(name + ".bias")
        relu = int(k["activation"] == "relu")
        if rank == 1 and k["data_format"] == "channels_last":
            B, L, C = x.shape
            Lo = c_oracle.qoc_conv1d_out_len(L, kern.shape[0], k["strides"][0], k["dilation_rate"][0], PADS[k["padding"]])
            y = np.empty((B, Lo, 4 * filters), np.float32)
            rc = c_oracle.qoc_conv1d_forward(_fp(x), _fp(kern), _fp(bias), _fp(y), B, L, C // 4, filters, kern.shape[0],
                                             k["strides"][0], k["dilation_rate"][0], PADS[k["padding"]], relu)
            assert rc == Lo
        elif rank == 2 and k["data_format"] == "channels_first":
            B, C, H, W = x.shape
            ref_shape = g[name + ".y"].shape
            y = np.empty(ref_shape, np.float32)
            rc = c_oracle.qoc_conv2d_cf_forward(_fp(x), _fp(kern), _fp(bias), _fp(y), B, H, W, C // 4, filters, kern.shape[0],
                                                kern.shape[1], k["strides"][0], k["strides"][1], k["dilation_rate"][0],
                                                k["dilation_rate"][1], PADS[k["padding"]], relu)
            assert rc == 0
        else:
            continue
        np.testing.assert_allclose(y, g[name + ".y"], rtol=1e-5, atol=1e-5, err_msg=name)
        checked += 1
    assert checked >= 12
    gd = golden.load("dense_forward")
    for name, xs, units, kw in DENSE_CASES:
        if kw.get("activation") not in (None, "relu"):
            continue
        x, kern, bias = gd[name + ".x"], gd[name + ".kernel"], gd.get(name + ".bias")
        y = np.empty((x.shape[0], units), np.float32)
        assert c_oracle.qoc_dense_forward(_fp(x), _fp(kern), _fp(bias), _fp(y), x.shape[0], x.shape[1] // 4, units // 4,
                                          int(kw.get("activation") == "relu")) == 0
        np.testing.assert_allclose(y, gd[name + ".y"], rtol=1e-5, atol=1e-5, err_msg=name)


def test_c_oracle_known_answers(c_oracle, golden):
    """(1+2i+3j+4k) (x) (5+6i+7j+8k) = -60+12i+30j+24k for the convolution, conj(w) (x) x = 70+0i-16j-8k for the dense layer."""
    kat = golden.load("kat")
    x = np.array([5, 6, 7, 8], np.float32)
    w = np.array([1, 2, 3, 4], np.float32)
    y = np.empty(4, np.float32)
    assert c_oracle.qoc_conv1d_forward(_fp(x), _fp(w), None, _fp(y), 1, 1, 1, 1, 1, 1, 1, 0, 0) == 1
    np.testing.assert_array_equal(y, kat["conv_w1234_x5678"].ravel())
    assert c_oracle.qoc_dense_forward(_fp(x), _fp(w), None, _fp(y), 1, 1, 1, 0) == 0
    np.testing.assert_array_equal(y, kat["dense_w1234_x5678"].ravel())
