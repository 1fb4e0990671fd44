"""Two-GPU test of the gradient path (skipped on a single-GPU box; run with `gpurun --gpus 2 -- pytest -m gpu
tests/test_gpu_multi.py`): each rank runs forward + backward of its batch shard through the product layers, the
kernel / bias gradients land directly in one flat device bucket, qnn_allreduce_f32 (NCCL over NVLink) sums it, and
the result equals the oracle's full-batch gradients."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    rng = np.random.default_rng(7)
    x = rng.normal(size=(12, 64, 32)).astype(np.float32)
    dy = rng.normal(size=(12, 64)).astype(np.float32)
    kc = (rng.normal(size=(3, 8, 64)) * 0.2).astype(np.float32)
    bc = (rng.normal(size=64) * 0.1).astype(np.float32)
    kd = (rng.normal(size=(1024, 64)) * 0.03).astype(np.float32)           # flatten(64 * 64) = 4096 = 4 * 1024
    bd = (rng.normal(size=64) * 0.1).astype(np.float32)
    return x, dy, kc, bc, kd, bd


def _worker(rank, world, port, out):
    import sys
    from conftest import PKG, REPO
    for p in (REPO, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["QNN_ALGO"] = "general"       # fp32 forward so that relu masks match the oracle exactly
    os.environ["QNN_MATH"] = "fp32"
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from complexnn import QuaternionConv1D, QuaternionDense
    from complexnn.dataparallel import GradBucket, allreduce_, destroy_comm, init_comm, shard_bounds
    x, dy, kc, bc, kd, bd = _problem()
    conv = QuaternionConv1D(16, 3, padding="same", activation="relu")
    conv.build((None, 64, 32))
    conv.built = True
    conv.set_weights([kc, bc])
    dense = QuaternionDense(64, activation="relu")
    dense.build((None, 4096))
    dense.built = True
    dense.set_weights([kd, bd])
    init_comm(rank, world)
    bucket = GradBucket([conv, dense], device="cuda")
    lo, hi = shard_bounds(x.shape[0], rank, world)
    xs, dys = torch.from_numpy(x[lo:hi]).cuda(), torch.from_numpy(dy[lo:hi]).cuda()
    h = conv(xs)
    hf = h.reshape(h.shape[0], -1)
    z = dense(hf)
    dkd, dbd = bucket.views(dense)
    dh, _, _ = dense.backward(hf, z, dys, grad_kernel_out=dkd, grad_bias_out=dbd)
    dkc, dbc = bucket.views(conv)
    conv.backward(xs, h, dh.reshape(h.shape), need_input_grad=False, grad_kernel_out=dkc, grad_bias_out=dbc)
    allreduce_(bucket)
    torch.cuda.synchronize()
    out[rank] = bucket.flat.cpu().numpy()
    destroy_comm()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_bucket_allreduce_matches_full_batch_oracle(native_lib):
    import torch.multiprocessing as mp
    from oracle import qoracle as O
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    x, dy, kc, bc, kd, bd = _problem()
    h = O.qconv_forward(x, kc, bc, 16, 1, "same", "channels_last", 1, "relu", out_dtype=None)
    dh, dkd, dbd = O.qdense_backward(h.reshape(12, -1), kd, bd, 64, "relu", dy)
    _, dkc, dbc = O.qconv_backward(x, kc, bc, 16, 1, "same", "channels_last", 1, "relu", dh.reshape(h.shape))
    full = np.concatenate([dkc.ravel(), dbc.ravel(), dkd.ravel(), dbd.ravel()])
    for r in range(world):
        err = np.abs(out[r] - full).max() / np.abs(full).max()
        assert err < 1e-4, err
    np.testing.assert_array_equal(out[0], out[1])
