"""Batch-sharded data parallelism around the quaternion layers (absent from the reference).

Forward: every rank runs its contiguous shard of the batch with replicated weights -- no collective.
Backward: each layer's `backward` writes its kernel / bias gradients straight into one flat fp32 bucket, which is then
summed over ranks in a single all-reduce: NCCL through the library's own communicator (`qnn_allreduce_f32`) for device
buckets, or `torch.distributed` (gloo) for host buckets in the CPU tests.
"""
import ctypes

import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous, balanced shard [lo, hi) of n units for `rank` of `world` (first n % world ranks get one more)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("need 0 <= rank < world")
    base, extra = divmod(int(n), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class GradBucket(object):
    """One flat fp32 buffer holding [dkernel_0, dbias_0, dkernel_1, ...] in layer order, with shaped views."""

    def __init__(self, layers, device="cpu"):
        import torch
        self.slots = []
        offset = 0
        for layer in layers:
            for var in layer.weights:
                n = int(np.prod(var.shape))
                self.slots.append((layer, var, offset, tuple(var.shape)))
                offset += n
        self.flat = torch.zeros(offset, dtype=torch.float32, device=device)

    def views(self, layer):
        """(dkernel_view, dbias_view | None) of `layer` inside the flat buffer."""
        out = [self.flat[o:o + int(np.prod(s))].view(s) for (l, _, o, s) in self.slots if l is layer]
        return out[0], (out[-1] if len(out) > 1 else None)

    def numel(self):
        return int(self.flat.numel())


_comm_ready = False


def init_comm(rank, world):
    """Create the library's NCCL communicator; the 128-byte unique id travels through torch.distributed."""
    global _comm_ready
    import torch
    import torch.distributed as dist
    from . import _native
    lib = _native.lib()
    uid = (ctypes.c_char * 128)()
    if rank == 0:
        _native.check(lib.qnn_comm_unique_id(ctypes.cast(uid, ctypes.c_void_p)))
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    buf = ctypes.create_string_buffer(raw, 128)
    _native.check(lib.qnn_comm_init(rank, world, ctypes.cast(buf, ctypes.c_void_p)))
    _comm_ready = True


def allreduce_(bucket, average_over=None):
    """In-place sum of the bucket over all ranks (then optional division)."""
    import torch
    flat = bucket.flat if isinstance(bucket, GradBucket) else bucket
    if flat.is_cuda and _comm_ready:
        from . import _native
        _native.check(_native.lib().qnn_allreduce_f32(ctypes.c_void_p(flat.data_ptr()), flat.numel(),
                                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    else:
        import torch.distributed as dist
        dist.all_reduce(flat)
    if average_over:
        flat.div_(average_over)
    return flat


def destroy_comm():
    global _comm_ready
    if _comm_ready:
        from . import _native
        _native.check(_native.lib().qnn_comm_destroy())
        _comm_ready = False
