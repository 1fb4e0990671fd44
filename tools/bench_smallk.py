#!/usr/bin/env python
"""Times the small-K (in_q < 4) forward against the general CUDA-core kernel on the reference's own small-K layers:
the first DECODA layer (models/example_model.py:25: QuaternionConv1D(32, 3, same, relu) on x[325, 250, 4]) and a
batch-scaled copy of it.  CUDA events over a graph of 50 launches; prints one JSON line per shape."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
sys.path[:0] = [REPO, PKG]
from complexnn import _ops  # noqa: E402
from complexnn._layer import Variable  # noqa: E402


def timed(fn, n=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.cuda.graph(g, stream=s):
        for _ in range(n):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


for (B, T, in_q, F, k) in ((325, 250, 1, 32, 3), (5200, 250, 1, 32, 3), (325, 250, 3, 64, 3)):
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.normal(size=(B, T, 4 * in_q)).astype(np.float32)).cuda()
    kern = Variable((rng.normal(size=(k, in_q, 4 * F)) * 0.3).astype(np.float32))
    bias = Variable(rng.normal(0, 0.1, 4 * F).astype(np.float32))
    run = lambda algo: _ops.conv_forward(x, kern, bias, F, (k,), (1,), "same", "channels_last", (1,), "relu", math="fp32",
                                         algo=algo)
    a, g = run("auto"), run("general")
    err = float((a - g).abs().max() / g.abs().max())
    t_small, t_gen = timed(lambda: run("auto")), timed(lambda: run("general"))
    by = 4 * (x.numel() + a.numel())
    print(json.dumps({"shape": "x[%d,%d,4x%d] -> %d filters, k=%d, same, relu" % (B, T, in_q, F, k),
                      "small_k_us": t_small, "general_us": t_gen, "speedup": t_gen / t_small,
                      "algorithmic_MB": by / 1e6, "small_k_GBps": by / t_small / 1e3, "max_rel_diff_vs_general": err}))

# The reference's OTHER small-K layer: the first TIMIT layer (models/interspeech_model.py:97), QuaternionConv2D(sf = 32,
# (3, 5), same, channels_first) on x[B, 4, 41, T] -- one quaternion input channel, 15 taps.  On CUDA cores it is compute-bound
# (7 680 FMA per output position vs 512 bytes of y); the streamed-sub-filter tensor-core kernel takes it after a
# channel-padding pre-pass (in_q 1 -> 8, zero channels).
for (B, T) in ((8, 400), (32, 400)):
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.normal(size=(B, 4, 41, T)).astype(np.float32)).cuda()
    kern = Variable((rng.normal(size=(3, 5, 1, 128)) * 0.3).astype(np.float32))
    bias = Variable(rng.normal(0, 0.1, 128).astype(np.float32))
    run = lambda algo, math: _ops.conv_forward(x, kern, bias, 32, (3, 5), (1, 1), "same", "channels_first", (1, 1), "linear",
                                               math=math, algo=algo)
    a, g = run("auto", "tf32"), run("general", "fp32")
    err = float((a - g).abs().max() / g.abs().max())
    t_tc, t_gen = timed(lambda: run("auto", "tf32"), 20), timed(lambda: run("general", "fp32"), 20)
    print(json.dumps({"shape": "TIMIT first layer x[%d,4,41,%d] channels_first -> 32 filters (3,5) same" % (B, T),
                      "tensor_core_padded_us": t_tc, "general_us": t_gen, "speedup": t_gen / t_tc,
                      "max_rel_diff_vs_general": err}))
