#!/usr/bin/env python
"""Random-shape parity fuzz on the GPU: the kernel the library picks (algo auto; TF32 and 3xTF32) against its own IEEE fp32
CUDA-core kernel on the same inputs -- a fast way to cover kernel-selection corners (split last round, ragged rows in place,
several filter-tile passes, starved layers) that the fixed test lists do not enumerate.  python tools/fuzz_parity.py [N] [seed]
Prints one summary JSON line; exit code 1 on any violation."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
sys.path[:0] = [REPO, PKG]
from complexnn import _native, _ops  # noqa: E402
from complexnn._layer import Variable  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
lib = _native.lib()
hist, worst, bad = {}, {"tf32": 0.0, "3xtf32": 0.0}, []
for it in range(N):
    rank = int(rng.choice([1, 1, 2]))
    layout = "channels_last" if rank == 1 or rng.random() < 0.6 else "channels_first"
    in_q = int(rng.choice([4, 8, 9, 12, 16, 24, 40, 41, 43, 44, 48, 64, 100]))
    F = int(rng.choice([16, 32, 48, 64, 96, 128, 192]))
    k = tuple(int(v) for v in rng.integers(1, 5, size=rank))
    d = tuple(int(v) for v in rng.integers(1, 3, size=rank))
    pad = str(rng.choice(["same", "valid"] + (["causal"] if rank == 1 else [])))
    if rank == 1:
        sp = (int(rng.choice([5, 64, 128, 131, 300, 700])),)
        B = int(rng.choice([1, 3, 40, 150, 300]))
    else:
        sp = (int(rng.choice([1, 3, 9])), int(rng.choice([8, 64, 132, 260])))
        B = int(rng.choice([1, 2, 20]))
    if pad == "valid":
        sp = tuple(max(n, (kk - 1) * dd + 2) for n, kk, dd in zip(sp, k, d))
    if layout == "channels_first":
        sp = sp[:-1] + ((sp[-1] + 3) // 4 * 4,)
    act = str(rng.choice(["relu", "linear", "tanh"]))
    shape = (B, 4 * in_q) + sp if layout == "channels_first" else (B,) + sp + (4 * in_q,)
    x = torch.from_numpy(rng.normal(size=shape).astype(np.float32)).cuda()
    kern = Variable((rng.normal(size=k + (in_q, 4 * F)) / np.sqrt(4 * in_q * np.prod(k))).astype(np.float32))
    bias = Variable(rng.normal(0, 0.1, 4 * F).astype(np.float32)) if rng.random() < 0.7 else None
    ones = (1,) * rank
    ref = _ops.conv_forward(x, kern, bias, F, k, ones, pad, layout, d, act, math="fp32", algo="general")
    scale = float(ref.abs().max()) + 1e-30
    for math, tol in (("tf32", 4e-3), ("3xtf32", 1e-4)):   # tf32: max over millions of outputs of short contractions reaches ~2e-3
        desc = _native.make_conv_desc(rank, B, sp, in_q, F, k, ones, d, pad, layout, act, math=math)
        kid = lib.qnn_conv_forward_kernel(ctypes.byref(desc))
        hist[kid] = hist.get(kid, 0) + 1
        y = _ops.conv_forward(x, kern, bias, F, k, ones, pad, layout, d, act, math=math, algo="auto")
        err = float((y - ref).abs().max()) / scale
        worst[math] = max(worst[math], err)
        if not err <= tol:
            bad.append({"shape": [rank, layout, B, list(sp), in_q, F, list(k), list(d), pad, act], "math": math, "kernel": kid,
                        "err": err})
    # gradients (relu / linear layers): the kernels the library picks vs the CUDA-core kernels, all on the fp32 forward's y
    if act in ("relu", "linear"):
        dy = torch.from_numpy(rng.normal(size=tuple(ref.shape)).astype(np.float32)).cuda()
        g_ref = _ops.conv_backward(x, ref, dy, kern, bias is not None, F, k, ones, pad, layout, d, act, math="fp32", algo="general")
        g_tc = _ops.conv_backward(x, ref, dy, kern, bias is not None, F, k, ones, pad, layout, d, act, math="tf32", algo="auto")
        for name, a, b in zip(("dx", "dkernel", "dbias"), g_tc, g_ref):
            if a is None:
                continue
            err = float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30)
            worst["grad_" + name] = max(worst.get("grad_" + name, 0.0), err)
            if not err <= 4e-3:
                bad.append({"shape": [rank, layout, B, list(sp), in_q, F, list(k), list(d), pad, act], "grad": name, "err": err})
torch.cuda.synchronize()
print(json.dumps({"shapes": N, "kernel_histogram (0 general, 1 resident, 2 streamed, 3 small-K)": hist, "worst_max_rel": worst,
                  "violations": bad}))
sys.exit(1 if bad else 0)
