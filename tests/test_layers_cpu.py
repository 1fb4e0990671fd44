"""CPU tests of the host side: the `complexnn` mirror keeps the reference's API surface (names, constructor
arguments, stored-weight layout, output shapes, config keys, error behaviour) and the C-ABI library loads and exports
what include/qnn.h declares.  No compute call is made here (that needs a GPU; see test_gpu_parity.py)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import complexnn
from complexnn import (QuaternionConv, QuaternionConv1D, QuaternionConv2D, QuaternionConv3D, QuaternionDense,
                       qconv_init, qdense_init, sqrt_init)
from cases import CONV_CASES, conv_kwargs
from oracle import qoracle as O

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")


def test_export_list_matches_reference():
    # reference complexnn/__init__.py:9-17
    names = ["QuaternionConv", "QuaternionConv1D", "QuaternionConv2D", "QuaternionConv3D", "QuaternionDense",
             "sqrt_init", "qdense_init", "qconv_init", "GetRFirst", "GetIFirst", "GetJFirst", "GetKFirst",
             "getpart_quaternion_output_shape_first", "get_rpart_first", "get_ipart_first", "get_jpart_first",
             "get_kpart_first"]
    for n in names:
        assert hasattr(complexnn, n), n
    from complexnn.conv import QuaternionConvolution1D, QuaternionConvolution2D, QuaternionConvolution3D  # conv.py:818-820
    assert QuaternionConvolution1D is QuaternionConv1D and QuaternionConvolution3D is QuaternionConv3D


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), os.path.join(root, f)
                assert "qoracle" not in src, os.path.join(root, f)


def test_initialisers_bit_exact_with_reference(golden):
    g = golden.load("init")
    np.random.seed(7)
    np.testing.assert_array_equal(qconv_init((3,), 5, 1, 6, "he")((3, 5, 6)), g["conv1d_he"])
    np.random.seed(8)
    np.testing.assert_array_equal(qconv_init((2, 3), 4, 2, 3, "glorot")((2, 3, 4, 3)), g["conv2d_glorot"])
    np.random.seed(9)
    np.testing.assert_array_equal(qdense_init((6, 5), "he")((6, 20)), g["dense_he"])
    np.random.seed(10)
    np.testing.assert_array_equal(qdense_init((4, 7), "glorot")((4, 28)), g["dense_glorot"])
    with pytest.raises(ValueError, match="Invalid criterion"):
        qdense_init((4, 7), "lecun")((4, 28))
    assert np.allclose(sqrt_init()((3,)), 1 / np.sqrt(2))


def test_stored_weight_layout():
    c = QuaternionConv1D(64, 3, padding="same", activation="relu")
    c.build((None, 256, 160))
    assert c.kernel_shape == (3, 40, 64)                 # what build declares (conv.py:165) ...
    assert c.kernel.shape == (3, 40, 256)                # ... and what the initialiser really returns
    assert c.bias.shape == (256,)
    assert [w.shape for w in c.weights] == [(3, 40, 256), (256,)]
    c2 = QuaternionConv2D(128, (3, 3), data_format="channels_first")
    c2.build((None, 256, 128, 128))
    assert c2.kernel.shape == (3, 3, 64, 512)
    d = QuaternionDense(256)                             # units = REAL outputs: 64 quaternion units
    d.build((None, 160))
    assert d.q_units == 64 and d.kernel.shape == (40, 256) and d.bias.shape == (256,)
    n = QuaternionConv(1, 4, 3, normalize_weight=True)   # gammas exist (and are never used), order as conv.py:175-278
    n.build((None, 10, 8))
    assert [w.name.split("/")[1] for w in n.weights] == ["kernel", "gamma_rr", "gamma_ri", "gamma_rj", "gamma_rk",
                                                          "gamma_ii", "gamma_ij", "gamma_ik", "gamma_jj", "gamma_jk",
                                                          "gamma_kk", "bias"]
    assert n.gamma_rr.shape == (2 * 4,) and np.allclose(n.gamma_rr.numpy(), 1 / np.sqrt(2)) and not n.gamma_ri.numpy().any()
    # the reference's own (odd) table, conv.py:238-251: gamma_jk gets the *diag* initialiser, gamma_kk the *off* one
    assert np.allclose(n.gamma_jk.numpy(), 1 / np.sqrt(2)) and not n.gamma_kk.numpy().any()
    assert np.allclose(n.gamma_ii.numpy(), 1 / np.sqrt(2)) and np.allclose(n.gamma_jj.numpy(), 1 / np.sqrt(2))
    nb = QuaternionDense(8, use_bias=False)
    nb.build((None, 12))
    assert nb.bias is None and len(nb.weights) == 1


def test_constructor_defaults_and_normalisation():
    c = QuaternionConv2D(8, 3)
    assert c.kernel_size == (3, 3) and c.strides == (1, 1) and c.dilation_rate == (1, 1)
    assert c.padding == "valid" and c.data_format == "channels_last" and c.use_bias and c.init_criterion == "he"
    assert c.activation.name == "linear" and c.input_spec.ndim == 4
    assert QuaternionConv3D(2, (1, 2, 3), strides=2).strides == (2, 2, 2)
    with pytest.raises(ValueError):
        QuaternionConv1D(8, (3, 3))
    with pytest.raises(ValueError):
        QuaternionConv1D(8, 3, padding="full")
    with pytest.raises(ValueError):
        QuaternionConv1D(8, 3, data_format="nchw")
    with pytest.raises(TypeError):
        QuaternionConv1D(8, 3, bogus=1)
    d = QuaternionDense(16, input_dim=8, name="head")
    assert d.name == "head" and d.batch_input_shape == (None, 8) and d.supports_masking
    assert QuaternionConv1D(1, 1).name.startswith("quaternion_conv1d_")


def test_error_behaviour_matches_reference():
    with pytest.raises(KeyError):                        # conv.py:167: only 'quaternion' is a key
        QuaternionConv1D(4, 3, kernel_initializer="quaternion_independent").build((None, 10, 8))
    with pytest.raises(ValueError, match="channel dimension"):       # conv.py:161-163
        QuaternionConv1D(4, 3).build((None, 10, None))
    with pytest.raises(AssertionError):                  # dense.py:94
        QuaternionDense(8).build((None, 4, 8))
    with pytest.raises(AssertionError):                  # dense.py:95
        QuaternionDense(8).build((None, 7))
    with pytest.raises(ValueError, match="ndim"):        # InputSpec(ndim=rank+2)
        QuaternionConv1D(4, 3)(np.zeros((2, 8), np.float32))
    d = QuaternionDense(8)
    d.build((None, 12))
    d.built = True
    with pytest.raises(ValueError, match="axis"):        # InputSpec(axes={-1: 4*in_q}) after build
        d(np.zeros((2, 16), np.float32))
    QuaternionDense(8, kernel_initializer="random_uniform").build((None, 8))    # silently ignored


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_compute_output_shape(golden, case):
    name, rank, xs, filters, ksz, kw = case
    cls = {1: QuaternionConv1D, 2: QuaternionConv2D, 3: QuaternionConv3D}[rank]
    layer = cls(filters, ksz, **kw)
    assert tuple(layer.compute_output_shape((None,) + xs[1:]))[1:] == golden.load("conv_forward")[name + ".y"].shape[1:]
    assert QuaternionDense(12).compute_output_shape((5, 8)) == (5, 12)


CONV_KEYS = {"filters", "kernel_size", "strides", "padding", "data_format", "dilation_rate", "activation", "use_bias",
             "normalize_weight", "kernel_initializer", "bias_initializer", "gamma_diag_initializer",
             "gamma_off_initializer", "kernel_regularizer", "bias_regularizer", "gamma_diag_regularizer",
             "gamma_off_regularizer", "activity_regularizer", "kernel_constraint", "bias_constraint",
             "gamma_diag_constraint", "gamma_off_constraint", "init_criterion", "spectral_parametrization", "rank"}
DENSE_KEYS = {"units", "activation", "use_bias", "init_criterion", "kernel_initializer", "bias_initializer",
              "kernel_regularizer", "bias_regularizer", "activity_regularizer", "kernel_constraint", "bias_constraint",
              "seed"}


def test_get_config_key_sets_and_json_safety():
    base = {"name", "trainable", "dtype"}
    cfg = QuaternionConv(2, 4, 3).get_config()                          # conv.py:375-401
    assert set(cfg) - base == CONV_KEYS
    cfg1 = QuaternionConv1D(4, 3, activation="relu").get_config()       # pops rank and data_format (conv.py:520-524)
    assert set(cfg1) - base == CONV_KEYS - {"rank", "data_format"}
    cfg2 = QuaternionConv2D(4, 3).get_config()                          # pops rank (conv.py:655-658)
    assert set(cfg2) - base == CONV_KEYS - {"rank"}
    assert set(QuaternionConv3D(4, 3).get_config()) - base == CONV_KEYS - {"rank"}
    d = QuaternionDense(8, activation="relu", seed=3)
    cfgd = d.get_config()                                               # dense.py:178-191
    assert set(cfgd) - base == DENSE_KEYS and cfgd["seed"] == 3
    for c in (cfg, cfg1, cfg2, cfgd):
        json.dumps(c)                                                   # the shipped reference fails here
    assert cfg1["activation"] == "relu" and cfg1["kernel_initializer"] == "quaternion"
    assert cfg1["gamma_diag_initializer"] == "sqrt_init" and cfg1["bias_initializer"]["class_name"] == "Zeros"
    clone = QuaternionConv1D.from_config(cfg1)
    assert clone.get_config() == cfg1
    clone_d = QuaternionDense.from_config(cfgd)
    assert clone_d.get_config() == cfgd


def test_get_set_weights_roundtrip():
    np.random.seed(0)
    c = QuaternionConv1D(4, 3)
    c.build((None, 9, 8))
    w = c.get_weights()
    assert [a.dtype for a in w] == [np.float32, np.float32] and w[0].shape == (3, 2, 16)
    new = [np.full_like(w[0], 0.5), np.arange(16, dtype=np.float32)]
    c.set_weights(new)
    np.testing.assert_array_equal(c.get_weights()[1], new[1])
    with pytest.raises(ValueError):
        c.set_weights([new[0]])
    with pytest.raises(ValueError):
        c.set_weights([new[0][:, :, :4], new[1]])
    assert c.count_params() == 3 * 2 * 16 + 16


def test_component_getters_blocked_layout():
    from complexnn import (get_rpart_first, get_ipart_first, get_jpart_first, get_kpart_first, GetKFirst,
                           getpart_quaternion_output_shape_first)
    x3 = np.arange(2 * 3 * 8).reshape(2, 3, 8)            # ndim 3 -> last axis (utils.py:25-27)
    np.testing.assert_array_equal(get_ipart_first(x3), x3[:, :, 2:4])
    x4 = np.arange(2 * 8 * 3 * 3).reshape(2, 8, 3, 3)     # otherwise axis 1 (channels_first hard-wired, utils.py:20-23)
    np.testing.assert_array_equal(get_rpart_first(x4), x4[:, 0:2])
    np.testing.assert_array_equal(get_jpart_first(x4), x4[:, 4:6])
    np.testing.assert_array_equal(get_kpart_first(x4), x4[:, 6:])
    x2 = np.arange(2 * 8).reshape(2, 8)
    np.testing.assert_array_equal(GetKFirst()(x2), x2[:, 6:])
    assert getpart_quaternion_output_shape_first((None, 8, 3, 3)) == (None, 2, 3, 3)
    assert getpart_quaternion_output_shape_first((None, 5, 8)) == (None, 5, 2)


# --------------------------------------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(REPO, "include", "qnn.h")).read()
    declared = set(re.findall(r"QNN_API[^;]*?\b(qnn_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 16
    from complexnn import _native
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    raw = ctypes.CDLL(_native.LIB_PATH)
    for sym in declared:
        assert hasattr(raw, sym), sym
    assert native_lib.qnn_abi_version() == 2
    assert native_lib.qnn_launch_count() == 0


def test_abi_geometry_matches_oracle_padding_rules(native_lib):
    from complexnn import _native
    rng = np.random.default_rng(0)
    for _ in range(300):
        rank = int(rng.integers(1, 4))
        sp = [int(rng.integers(0, 30)) for _ in range(rank)]
        ks = [int(rng.integers(1, 6)) for _ in range(rank)]
        st = [int(rng.integers(1, 4)) for _ in range(rank)]
        dl = [int(rng.integers(1, 4)) for _ in range(rank)]
        pad = ["valid", "same", "causal"][int(rng.integers(0, 3 if rank == 1 else 2))]
        d = _native.make_conv_desc(rank, 2, sp, 4, 8, ks, st, dl, pad, "channels_last", "relu")
        out = (ctypes.c_int32 * 3)()
        assert native_lib.qnn_conv_out_spatial(ctypes.byref(d), ctypes.byref(out)) == 0
        expect = [O.pad_amounts(sp[a], ks[a], st[a], dl[a], pad)[2] for a in range(rank)]
        assert list(out)[:rank] == expect, (sp, ks, st, dl, pad)
        assert expect == [max(O.conv_output_length(sp[a], ks[a], pad, st[a], dl[a]), 0) for a in range(rank)]


def test_abi_argument_errors_and_kernel_selection(native_lib):
    from complexnn import _native
    bad = _native.make_conv_desc(2, 1, (4, 4), 4, 16, (3, 3), (1, 1), (1, 1), "causal", "channels_last", None)
    out = (ctypes.c_int32 * 3)()
    assert native_lib.qnn_conv_out_spatial(ctypes.byref(bad), ctypes.byref(out)) == -1
    assert b"causal" in native_lib.qnn_last_error()
    with pytest.raises(ValueError, match="causal"):
        _native.check(-1)
    bad2 = _native.make_conv_desc(1, 1, (4,), 0, 16, (3,), (1,), (1,), "same", "channels_last", None)
    assert native_lib.qnn_conv_out_spatial(ctypes.byref(bad2), ctypes.byref(out)) == -1
    # BASELINE config 2 and the north-star dense shape run on the tensor-core kernel, ragged channel counts do not
    cfg2 = _native.make_conv_desc(1, 256, (256,), 40, 64, (3,), (1,), (1,), "same", "channels_last", "relu")
    assert native_lib.qnn_conv_uses_tensor_cores(ctypes.byref(cfg2)) == 1
    assert native_lib.qnn_dense_uses_tensor_cores(65536, 40, 64) == 1
    assert native_lib.qnn_dense_uses_tensor_cores(32, 250, 128) == 1          # DECODA first layer: in_q % 4 != 0 -> ragged stage mode
    assert native_lib.qnn_dense_uses_tensor_cores(32, 250, 100) == 0          # q_units % 16 != 0
    tim = _native.make_conv_desc(1, 256, (256,), 41, 64, (3,), (1,), (1,), "same", "channels_last", "relu")
    assert native_lib.qnn_conv_uses_tensor_cores(ctypes.byref(tim)) == 1      # cfg 3 first layer (TIMIT, in_q = 41)
    s2 = _native.make_conv_desc(1, 8, (64,), 40, 64, (3,), (2,), (1,), "same", "channels_last", "relu")
    assert native_lib.qnn_conv_uses_tensor_cores(ctypes.byref(s2)) == 1          # strides 2..4: several row boxes per x stage
    s5 = _native.make_conv_desc(1, 8, (64,), 40, 64, (3,), (5,), (1,), "same", "channels_last", "relu")
    assert native_lib.qnn_conv_uses_tensor_cores(ctypes.byref(s5)) == 0
    s2cf = _native.make_conv_desc(2, 8, (16, 16), 64, 128, (3, 3), (2, 2), (1, 1), "same", "channels_first", "relu")
    cf = _native.make_conv_desc(2, 8, (16, 16), 64, 128, (3, 3), (1, 1), (1, 1), "same", "channels_first", "relu")
    assert native_lib.qnn_conv_uses_tensor_cores(ctypes.byref(cf)) == 1          # channels_first tensor-core kernel
    cfg5 = _native.make_conv_desc(2, 128, (128, 128), 64, 128, (3, 3), (1, 1), (1, 1), "same", "channels_first", "relu")
    assert native_lib.qnn_conv_uses_tensor_cores(ctypes.byref(cfg5)) == 1        # BASELINE config 5
    cf_odd = _native.make_conv_desc(2, 8, (16, 18), 64, 128, (3, 3), (1, 1), (1, 1), "same", "channels_first", "relu")
    assert native_lib.qnn_conv_uses_tensor_cores(ctypes.byref(cf_odd)) == 1      # row length % 4 != 0: row-padded scratch copies
    timit_1 = _native.make_conv_desc(2, 4, (41, 200), 1, 32, (3, 5), (1, 1), (1, 1), "same", "channels_first", "linear")
    assert native_lib.qnn_conv_forward_kernel(ctypes.byref(timit_1)) == _native.KERNEL_TC_CF   # interspeech_model.py:97: ONE
    #                                                  quaternion input channel, padded to 8 in a pre-pass for the tensor cores
    timit_t = _native.make_conv_desc(2, 4, (41, 333), 32, 32, (3, 5), (1, 1), (1, 1), "same", "channels_first", "linear")
    assert native_lib.qnn_conv_forward_kernel(ctypes.byref(timit_t)) == _native.KERNEL_TC_CF   # free time axis, T = 333
    cl2 = _native.make_conv_desc(2, 8, (16, 17), 64, 128, (3, 3), (1, 1), (1, 1), "same", "channels_last", "relu")
    assert native_lib.qnn_conv_forward_kernel(ctypes.byref(cl2)) == _native.KERNEL_TC_CF   # channels_last rank 2, any row length
    big = _native.make_conv_desc(1, 8, (300,), 128, 128, (5,), (1,), (1,), "same", "channels_last", "relu")
    assert native_lib.qnn_conv_forward_kernel(ctypes.byref(big)) == _native.KERNEL_TC_CF   # rank 1 whose sub-filters (1.3 MB)
    #                                                                                         cannot stay resident: streamed
    x3 = _native.make_conv_desc(1, 256, (256,), 64, 64, (3,), (1,), (1,), "same", "channels_last", "relu", math="3xtf32")
    assert native_lib.qnn_conv_forward_kernel(ctypes.byref(x3)) == _native.KERNEL_TC_CF    # cfg 3 layer under 3xTF32 (hi | lo
    #                                                                                         image 393 KB): streamed too
    # in_q < 4 (first DECODA layer, models/example_model.py:25): the small-K shuffle kernel in every math mode
    dec = _native.make_conv_desc(1, 325, (250,), 1, 32, (3,), (1,), (1,), "same", "channels_last", "relu")
    assert native_lib.qnn_conv_forward_kernel(ctypes.byref(dec)) == _native.KERNEL_SMALL_K
    assert native_lib.qnn_conv_uses_tensor_cores(ctypes.byref(dec)) == 0
    assert native_lib.qnn_conv_forward_kernel(ctypes.byref(cfg2)) == _native.KERNEL_TC_ROWS
    assert native_lib.qnn_conv_forward_kernel(ctypes.byref(cfg5)) == _native.KERNEL_TC_CF
    assert native_lib.qnn_conv_forward_kernel(ctypes.byref(s2)) == _native.KERNEL_TC_ROWS
    assert native_lib.qnn_conv_forward_kernel(ctypes.byref(s2cf)) == _native.KERNEL_GENERAL   # strided 2-D: general kernel
    dec.algo = _native.ALGO["general"]
    assert native_lib.qnn_conv_forward_kernel(ctypes.byref(dec)) == _native.KERNEL_GENERAL
    assert native_lib.qnn_allreduce_f32(None, 4, None) == -6                  # QNN_E_STATE: no communicator yet
    assert native_lib.qnn_comm_init(2, 2, None) == -1


def test_no_silent_cpu_fallback():
    """Without a GPU the layer call must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    layer = QuaternionDense(16)
    with pytest.raises(RuntimeError, match="qnn error"):
        layer(np.zeros((4, 16), np.float32))


def test_backward_kernel_selection_is_host_logic(native_lib):
    """qnn_*_backward_uses_tensor_cores: which gradients run on the tensor cores (no GPU needed to ask)."""
    from complexnn import _native

    def ask(desc):
        a, b = ctypes.c_int32(-1), ctypes.c_int32(-1)
        assert native_lib.qnn_conv_backward_uses_tensor_cores(ctypes.byref(desc), ctypes.byref(a), ctypes.byref(b)) == 0
        return a.value, b.value

    mk = _native.make_conv_desc
    # cfg 3 / 4 stack: inner conv layers (in_q = 64, F = 64): both gradients on tensor cores
    assert ask(mk(1, 256, (256,), 64, 64, (3,), (1,), (1,), "same", "channels_last", "relu")) == (1, 1)
    # first layer (TIMIT in_q = 41): kernel gradient yes (flat-row x stage); data gradient would need F' = 41 filters
    assert ask(mk(1, 256, (256,), 41, 64, (3,), (1,), (1,), "same", "channels_last", "relu")) == (0, 1)
    # cfg 2 (in_q = 40): dx needs a multiple of 16 "filters" in the transposed problem -> CUDA cores; dkernel on tensor cores
    assert ask(mk(1, 256, (256,), 40, 64, (3,), (1,), (1,), "same", "channels_last", "relu")) == (0, 1)
    # strided, fp32 math, general algo: CUDA-core kernels
    assert ask(mk(1, 8, (64,), 64, 64, (3,), (2,), (1,), "same", "channels_last", "relu")) == (0, 0)
    # channels_first rank 2 (what models/interspeech_model.py trains): the data gradient is the channels_first forward
    # kernel on dz with the transposed, tap-flipped image; the kernel gradient runs the position-contraction kernel once
    # per kernel row on channels_last scratch copies of x and dz
    assert ask(mk(2, 8, (16, 16), 64, 128, (3, 3), (1, 1), (1, 1), "same", "channels_first", "relu")) == (1, 1)
    assert ask(mk(2, 8, (16, 16), 64, 128, (3, 3), (1, 1), (1, 1), "same", "channels_last", "relu")) == (1, 1)
    # the TIMIT layers of models/interspeech_model.py:51-61,116: QuaternionConv2D(32, (3, 5), same, channels_first), in_q = 32
    assert ask(mk(2, 4, (41, 200), 32, 32, (3, 5), (1, 1), (1, 1), "same", "channels_first", "linear")) == (1, 1)
    assert ask(mk(2, 8, (16, 16), 64, 128, (3, 3), (1, 1), (1, 1), "same", "channels_first", "relu", math="3xtf32")) == (1, 1)
    # 3xTF32: both gradients on the tensor cores with three MMAs per block (data gradient: the streamed-sub-filter kernel
    # takes what does not fit resident -- a 3-tap conv with in_q = 64 needs a 393 KB hi | lo image; kernel gradient: hi | lo
    # A slots and dz stages)
    assert ask(mk(1, 256, (256,), 64, 64, (1,), (1,), (1,), "same", "channels_last", "relu", math="3xtf32")) == (1, 1)
    assert ask(mk(1, 256, (256,), 64, 64, (3,), (1,), (1,), "same", "channels_last", "relu", math="3xtf32")) == (1, 1)
    assert ask(mk(1, 8, (64,), 64, 64, (3,), (1,), (1,), "same", "channels_last", "relu", math="fp32")) == (0, 0)
    assert ask(mk(1, 8, (64,), 64, 64, (3,), (1,), (1,), "same", "channels_last", "relu", algo="general")) == (0, 0)
    # tanh has no fused derivative: the backward entry points refuse it, the query says "not on tensor cores"
    assert ask(mk(1, 8, (64,), 64, 64, (3,), (1,), (1,), "same", "channels_last", "tanh")) == (0, 0)
    a, b = ctypes.c_int32(-1), ctypes.c_int32(-1)
    assert native_lib.qnn_dense_backward_uses_tensor_cores(65536, 64, 64, ctypes.byref(a), ctypes.byref(b)) == 0
    assert (a.value, b.value) == (1, 1)
    assert native_lib.qnn_dense_backward_uses_tensor_cores(325, 250, 128, ctypes.byref(a), ctypes.byref(b)) == 0
    assert (a.value, b.value) == (0, 0)          # DECODA first layer: 250 "filters" do not tile in the transposed problem,
    #                                              and a 32-row x stage of 4 x 252 channels (129 KB) does not fit twice
    assert native_lib.qnn_dense_backward_uses_tensor_cores(325, 128, 128, ctypes.byref(a), ctypes.byref(b)) == 0
    assert (a.value, b.value) == (1, 1)          # DECODA QDNN layers 2 and 3
    assert native_lib.qnn_conv_backward_uses_tensor_cores(None, ctypes.byref(a), ctypes.byref(b)) == -1


def test_variable_assign_after_parameter_and_device_move():
    """set_weights after a training step (the mirror is then an autograd leaf) must not raise, and the
    newest values must survive; `version` moves with every update (it keys the packed-weight caches)."""
    import torch
    from complexnn._layer import Variable
    v = Variable(np.arange(6, dtype=np.float32).reshape(2, 3))
    p = v.parameter("cpu")
    assert p.requires_grad
    v0 = v.version
    v.assign(np.full((2, 3), 7.0, np.float32))               # used to raise: in-place copy into a leaf that requires grad
    assert v.version == v0 + 1
    np.testing.assert_array_equal(v.numpy(), np.full((2, 3), 7.0, np.float32))
    np.testing.assert_array_equal(p.detach().numpy(), np.full((2, 3), 7.0, np.float32))
    with torch.no_grad():
        p.add_(1.0)                                          # an optimiser step on the mirror
    v.mark_device_updated()
    assert v.version == v0 + 2
    np.testing.assert_array_equal(v.numpy(), np.full((2, 3), 8.0, np.float32))
    with pytest.raises(ValueError):
        v.assign(np.zeros((3, 2), np.float32))


def test_packed_image_queries_agree_with_kernel_selection(native_lib):
    """The packed kernel image a caller builds must be the one the forward will read: `*_packed_bytes` > 0 exactly when a
    tensor-core kernel takes the problem, for every activation and math mode -- and a dense layer's kernel must not depend
    on its activation (the dense packing ABI does not carry it; a rule that did once packed the streamed layout for a
    resident-kernel run)."""
    from complexnn import _native
    tc = (_native.KERNEL_TC_ROWS, _native.KERNEL_TC_CF)
    rng = np.random.default_rng(0)
    for _ in range(300):
        rank = int(rng.integers(1, 4))
        cf = bool(rng.integers(0, 2))
        in_q = int(rng.choice([1, 3, 4, 8, 16, 41, 64, 128]))
        F = int(rng.choice([8, 16, 32, 48, 64, 128]))
        k = tuple(int(v) for v in rng.integers(1, 6, size=rank))
        sp = tuple(int(v) for v in rng.choice([1, 7, 40, 131, 256], size=rank))
        act = str(rng.choice(["relu", "linear", "tanh"]))
        math = str(rng.choice(["tf32", "3xtf32", "fp32"]))
        pad = str(rng.choice(["same", "valid"]))
        d = _native.make_conv_desc(rank, 2, sp, in_q, F, k, (1,) * rank, (1,) * rank, pad,
                                   "channels_first" if cf else "channels_last", act, math=math)
        kern = native_lib.qnn_conv_forward_kernel(ctypes.byref(d))
        nbytes = native_lib.qnn_conv_packed_bytes(ctypes.byref(d), _native.PACK_FORWARD)
        assert (nbytes > 0) == (kern in tc), (rank, cf, in_q, F, k, sp, act, math, kern, nbytes)
        if kern in tc:
            per_part = int(np.prod(k)) * (-(-in_q // 8) * 8) * 4 * F * 4     # both images pad in_q to a multiple of 8
            assert nbytes == per_part * (2 if math == "3xtf32" else 1)
    for in_q, q_units in ((40, 64), (128, 128), (250, 128), (64, 16), (3, 64), (16, 100)):
        for math in ("tf32", "3xtf32", "fp32"):
            m = _native.MATH[math]
            kerns = {native_lib.qnn_dense_forward_kernel(1000, in_q, q_units, _native.ACT[a], m, 0)
                     for a in ("linear", "relu", "tanh", "selu")}
            assert len(kerns) == 1, (in_q, q_units, math, kerns)
            nbytes = native_lib.qnn_dense_packed_bytes(1000, in_q, q_units, m, 0, _native.PACK_FORWARD)
            assert (nbytes > 0) == (kerns.pop() in tc), (in_q, q_units, math)


def test_work_split_covers_every_item_exactly_once(native_lib):
    """Host logic of the persistent tensor-core forward kernels (qnn_conv_work_split, no GPU needed): the (grid, whole
    rounds, last-round items, split) plan together with the device-side mapping -- WorkItem in qnn_hamilton_tc.cu, Work in
    qnn_hamilton_tc2d.cu, restated below -- must compute every (item, filter half) exactly once, split only 64-wide
    filter tiles, and never use more CTAs than SMs."""
    from complexnn import _native
    n_sm = 148  # what the library assumes when no device is present
    out = (ctypes.c_int32 * 4)()

    def items_of(grid, full, rem, split):
        """(item, first filter, filters) of every CTA, as the kernels enumerate them"""
        got = []
        for cta in range(grid):
            n = full + (1 if cta < (2 * rem if split else rem) else 0)
            for k in range(n):
                if k < full or not split:
                    got.append((cta + k * grid, 0, 64 if split else None))
                else:
                    got.append((full * grid + (cta >> 1), (cta & 1) * 32, 32))
        return got

    cases = []
    # cfg 2: 512 tiles = 3 rounds + 68 -> 136 half items; cfg 2 with a full last round; small problems; dense
    for batch, T, in_q, F, k in ((256, 256, 40, 64, 3), (296, 128, 40, 64, 3), (10, 128, 40, 64, 3), (75, 128, 40, 64, 3),
                                 (256, 256, 8, 32, 3), (3, 1000, 16, 64, 5), (256, 256, 64, 64, 3), (300, 256, 64, 128, 3)):
        cases.append(_native.make_conv_desc(1, batch, (T,), in_q, F, (k,), (1,), (1,), "same", "channels_last", "relu"))
    cases.append(_native.make_conv_desc(2, 4, (30, 200), 64, 128, (3, 3), (1, 1), (1, 1), "same", "channels_first", "relu"))
    cases.append(_native.make_conv_desc(2, 1, (5, 131), 16, 64, (3, 5), (1, 1), (1, 1), "same", "channels_last", "relu"))
    seen_split = seen_whole = False
    for d in cases:
        kern = native_lib.qnn_conv_work_split(ctypes.byref(d), out)
        assert kern in (_native.KERNEL_TC_ROWS, _native.KERNEL_TC_CF), kern
        grid, full, rem, split = (int(v) for v in out)
        assert 1 <= grid <= n_sm and rem >= 0 and full >= 0
        total = full * grid + rem if full else rem
        assert (not full) or grid == n_sm
        if split:
            assert 0 < 2 * rem <= n_sm
        got = items_of(grid, full, rem, split)
        whole = sorted(i for i, f0, fe in got if fe != 32)
        halves = sorted((i, f0) for i, f0, fe in got if fe == 32)
        first_split = full * grid if split else total
        assert whole == list(range(first_split)), "whole items"
        assert halves == [(i, f0) for i in range(first_split, total) for f0 in (0, 32)], "half items"
        seen_split |= bool(split)
        seen_whole |= not split
    assert seen_split and seen_whole
    # a CUDA-core problem has no persistent grid
    d = _native.make_conv_desc(1, 4, (100,), 8, 24, (3,), (2,), (1,), "same", "channels_last", "relu")
    assert native_lib.qnn_conv_work_split(ctypes.byref(d), out) == _native.KERNEL_GENERAL and list(out) == [0, 0, 0, 0]
