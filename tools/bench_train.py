#!/usr/bin/env python
"""BASELINE.json configs[2] / [3]: the QCNN stack as BASELINE words it -- 3 x QuaternionConv1D(64, 3, same, relu) +
2 x QuaternionDense(256, relu) on TIMIT-shaped input x[B, 256, 4*41] -- forward only (`--fwd`) or forward + backward
with the kernel / bias gradients written into one flat bucket and summed over ranks with qnn_allreduce_f32 (NCCL).

  python tools/bench_train.py [--batch 256] [--steps 20] [--fwd]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py   (weak scaling)

Every rank runs the same per-GPU batch; time = max over ranks (CUDA events); qMACs are counted x1 (forward) or x3
(forward + data gradient + kernel gradient; the first layer has no data gradient -> x2).  Prints one JSON line."""
import argparse
import json
import os
import sys

if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":   # keep NCCL's version banner off stdout (one JSON line)
    os.environ["NCCL_DEBUG"] = "WARN"

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
sys.path[:0] = [REPO, PKG]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--T", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--fwd", action="store_true")
    ap.add_argument("--graph", action="store_true", help="capture one step in a CUDA graph and replay it (no phase split)")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import complexnn
    from complexnn import _native
    from complexnn.dataparallel import GradBucket, allreduce_, destroy_comm, init_comm
    np.random.seed(0)
    B, T = args.batch, args.T
    convs = [complexnn.QuaternionConv1D(64, 3, padding="same", activation="relu") for _ in range(3)]
    denses = [complexnn.QuaternionDense(256, activation="relu") for _ in range(2)]
    x = torch.randn(B, T, 164, device="cuda")
    h = x
    for l in convs:
        h = l(h)
    h = h.reshape(B * T, 256)
    for l in denses:
        h = l(h)
    layers = convs + denses
    if world > 1:
        init_comm(rank, world)
    bucket = GradBucket(layers, device="cuda")
    dy = torch.randn(B * T, 256, device="cuda") / (B * T)
    in_q = [41, 64, 64, 64, 64]
    taps = [3, 3, 3, 1, 1]
    q_fwd = sum(B * T * t * q * 64 for t, q in zip(taps, in_q))
    q_total = q_fwd if args.fwd else 3 * q_fwd - B * T * taps[0] * in_q[0] * 64

    def step(ev=None):
        acts = [x]
        h = x
        for l in convs:
            h = l(h)
            acts.append(h)
        h = h.reshape(B * T, 256)
        acts[-1] = h.view(B, T, 256)
        hs = [h]
        for l in denses:
            h = l(h)
            hs.append(h)
        if ev:
            ev[1].record()
        if args.fwd:
            return
        g = dy
        for i in (1, 0):
            dk, db = bucket.views(denses[i])
            g, _, _ = denses[i].backward(hs[i], hs[i + 1], g, grad_kernel_out=dk, grad_bias_out=db)
        g = g.view(B, T, 256)
        for i in (2, 1, 0):
            dk, db = bucket.views(convs[i])
            g, _, _ = convs[i].backward(acts[i], acts[i + 1], g, need_input_grad=i > 0, grad_kernel_out=dk, grad_bias_out=db)
        if ev:
            ev[2].record()
        if world > 1:
            allreduce_(bucket, average_over=world)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    graph = None
    if args.graph and world > 1:
        # Capturing the step INCLUDING the NCCL all-reduce of the library's own communicator hung on 8 GPUs (no error,
        # every rank stuck; round 1, cost the rest of the round's GPU budget).  Until that is understood the multi-GPU
        # step is launched eagerly.
        if rank == 0:
            print("--graph is single-GPU only for now (NCCL all-reduce inside the captured step hangs); running eagerly",
                  file=sys.stderr)
    elif args.graph:
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap):
                step()
        torch.cuda.current_stream().wait_stream(cap)
        graph.replay()
        torch.cuda.synchronize()
    l0 = _native.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_f = t_b = t_a = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        if graph is not None:
            graph.replay()
            continue
        ev[0].record()
        step(ev)
        ev[3].record()
        ev[3].synchronize()
        t_f += ev[0].elapsed_time(ev[1])
        if not args.fwd:
            t_b += ev[1].elapsed_time(ev[2])
            t_a += ev[2].elapsed_time(ev[3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = _native.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            "workload": "QCNN stack 3 x QConv1D(64,3,same,relu) + 2 x QDense(256,relu), x[%d,%d,164]%s" % (
                B, T, " forward" if args.fwd else " forward + backward + bucket all-reduce"),
            "n_gpus": world, "per_gpu_batch": B, "ms_per_step": ms, "qmacs_per_step_per_gpu": q_total,
            "value": world * q_total / (ms * 1e-3), "unit": "qMAC/s",
            "phases_ms": {"forward": t_f / args.steps, "backward": t_b / args.steps, "allreduce": t_a / args.steps},
            "launch": "CUDA graph replay of one captured step" if graph is not None else "eager, one C-ABI call per layer and pass",
            "bucket_floats": bucket.numel(), "launches_per_step": launches / args.steps, "steps": args.steps}))
    if world > 1:
        destroy_comm()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
