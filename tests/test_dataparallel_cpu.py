"""World-size-2 gloo test (CPU) of the data-parallel host logic: batch shards + one flat gradient bucket + a single
sum all-reduce reproduce the full-batch gradients.  The per-shard gradients come from the oracle (no GPU here); on the
GPU box the same bucket is filled by qnn_*_backward and reduced by qnn_allreduce_f32 (tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest

from complexnn import QuaternionConv1D, QuaternionDense
from complexnn.dataparallel import GradBucket, allreduce_, shard_bounds
from oracle import qoracle as O


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(10, 12, 8)).astype(np.float32)                     # conv input [B, T, 4*in_q]
    dy = rng.normal(size=(10, 16)).astype(np.float32)                       # gradient at the dense output
    kc = rng.normal(size=(3, 2, 16)).astype(np.float32) * 0.3
    bc = rng.normal(size=16).astype(np.float32) * 0.1
    kd = rng.normal(size=(48, 16)).astype(np.float32) * 0.1                 # flatten(12 * 16) = 192 = 4 * 48
    bd = rng.normal(size=16).astype(np.float32) * 0.1
    return x, dy, kc, bc, kd, bd


def _grads(x, dy, kc, bc, kd, bd):
    """conv1d(relu) -> flatten -> dense(relu): forward + backward with the oracle; returns flat [dkc, dbc, dkd, dbd]."""
    h = O.qconv_forward(x, kc, bc, 4, 1, "same", "channels_last", 1, "relu", out_dtype=None)
    hf = h.reshape(h.shape[0], -1)
    dh, dkd, dbd = O.qdense_backward(hf, kd, bd, 16, "relu", dy)
    _, dkc, dbc = O.qconv_backward(x, kc, bc, 4, 1, "same", "channels_last", 1, "relu", dh.reshape(h.shape))
    return np.concatenate([dkc.ravel(), dbc.ravel(), dkd.ravel(), dbd.ravel()])


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, dy, kc, bc, kd, bd = _problem()
    conv = QuaternionConv1D(4, 3, padding="same", activation="relu")
    conv.build((None, 12, 8))
    dense = QuaternionDense(16, activation="relu")
    dense.build((None, 192))
    bucket = GradBucket([conv, dense])
    assert bucket.numel() == kc.size + bc.size + kd.size + bd.size
    lo, hi = shard_bounds(x.shape[0], rank, world)
    bucket.flat.copy_(torch.from_numpy(_grads(x[lo:hi], dy[lo:hi], kc, bc, kd, bd).astype(np.float32)))
    dk_view, db_view = bucket.views(dense)
    assert tuple(dk_view.shape) == (48, 16) and tuple(db_view.shape) == (16,)
    assert dk_view.data_ptr() == bucket.flat[kc.size + bc.size:].data_ptr()     # views alias the flat buffer
    allreduce_(bucket)
    out[rank] = bucket.flat.numpy().copy()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_full_batch():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    full = _grads(*_problem())
    for r in range(world):
        np.testing.assert_allclose(out[r], full, rtol=2e-5, atol=2e-5)
    np.testing.assert_array_equal(out[0], out[1])
