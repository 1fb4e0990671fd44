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


def _worker(rank, world, port, out, mode="fp32"):
    import sys
    from conftest import PKG, REPO
    for p in (REPO, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if mode == "fp32":
        os.environ["QNN_ALGO"] = "general"       # fp32 forward so that relu masks match the oracle exactly
        os.environ["QNN_MATH"] = "fp32"
    else:
        os.environ["QNN_ALGO"] = "auto"          # tensor-core forward, data and kernel gradients into the bucket
        os.environ["QNN_MATH"] = "tf32"
    act = "relu" if mode == "fp32" else None     # (a tf32 forward flips relu masks next to zero: linear layers there)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from complexnn import QuaternionConv1D, QuaternionDense
    from complexnn.dataparallel import GradBucket, allreduce_, destroy_comm, init_comm, shard_bounds
    x, dy, kc, bc, kd, bd = _problem()
    conv = QuaternionConv1D(16, 3, padding="same", activation=act)
    conv.build((None, 64, 32))
    conv.built = True
    conv.set_weights([kc, bc])
    dense = QuaternionDense(64, activation=act)
    if mode == "fp32":
        dense.build((None, 4096))
        dense.built = True
        dense.set_weights([kd, bd])
    else:                                        # per-position dense (in_q = 16 -> 16 quaternion units): tensor-core shapes
        dense.build((None, 64))
        dense.built = True
        dense.set_weights([kd[:16], bd])
    init_comm(rank, world)
    bucket = GradBucket([conv, dense], device="cuda")
    lo, hi = shard_bounds(x.shape[0], rank, world)
    xs, dys = torch.from_numpy(x[lo:hi]).cuda(), torch.from_numpy(dy[lo:hi]).cuda()
    h = conv(xs)
    hf = h.reshape(h.shape[0], -1) if mode == "fp32" else h.reshape(-1, 64)
    if mode != "fp32":
        import ctypes
        from complexnn import _native
        a, b = ctypes.c_int32(-1), ctypes.c_int32(-1)
        assert _native.lib().qnn_dense_backward_uses_tensor_cores(hf.shape[0], 16, 16, ctypes.byref(a), ctypes.byref(b)) == 0
        assert (a.value, b.value) == (1, 1)
        dys = torch.from_numpy(np.repeat(dy[lo:hi], 64, axis=0)).cuda() / 64.0
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
@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_two_gpu_bucket_allreduce_matches_full_batch_oracle(native_lib, mode):
    import torch.multiprocessing as mp
    from oracle import qoracle as O
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, out, mode)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    x, dy, kc, bc, kd, bd = _problem()
    if mode == "fp32":
        h = O.qconv_forward(x, kc, bc, 16, 1, "same", "channels_last", 1, "relu", out_dtype=None)
        dh, dkd, dbd = O.qdense_backward(h.reshape(12, -1), kd, bd, 64, "relu", dy)
        _, dkc, dbc = O.qconv_backward(x, kc, bc, 16, 1, "same", "channels_last", 1, "relu", dh.reshape(h.shape))
        tol = 1e-4
    else:
        h = O.qconv_forward(x, kc, bc, 16, 1, "same", "channels_last", 1, None, out_dtype=None)
        dh, dkd, dbd = O.qdense_backward(h.reshape(-1, 64), kd[:16], bd, 64, None, np.repeat(dy, 64, axis=0) / 64.0)
        _, dkc, dbc = O.qconv_backward(x, kc, bc, 16, 1, "same", "channels_last", 1, None, dh.reshape(h.shape))
        tol = 2e-3                     # tf32 operands in the forward, the data gradient and the kernel gradients
    full = np.concatenate([dkc.ravel(), dbc.ravel(), dkd.ravel(), dbd.ravel()])
    for r in range(world):
        err = np.abs(out[r] - full).max() / np.abs(full).max()
        assert err < tol, err
    np.testing.assert_array_equal(out[0], out[1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_process_two_devices(native_lib):
    """library state is per device -- the same process drives cuda:0 then cuda:1 (tensor-core kernel with
    > 48 KB dynamic shared memory, stream-ordered scratch, host-buffer path) and gets identical results."""
    import sys
    from conftest import PKG, REPO
    for p in (REPO, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import complexnn
    rng = np.random.default_rng(3)
    x = rng.normal(size=(8, 200, 160)).astype(np.float32)
    np.random.seed(3)
    conv = complexnn.QuaternionConv1D(64, 3, padding="same", activation="relu")
    ys = []
    for d in (0, 1, 0):
        with torch.cuda.device(d):
            ys.append(conv(torch.from_numpy(x).to("cuda:%d" % d)).cpu().numpy())
            ys.append(conv(x))                                   # host-buffer path on the same device
    for y in ys[1:]:
        np.testing.assert_array_equal(ys[0], y)
