#!/usr/bin/env python
"""Times one channels_last QuaternionConv1D forward shape (math tf32, algo auto, packed image cached by the op layer) and
says which kernel the library picked:  python tools/bench_shape.py B T in_q F k [B T in_q F k ...]
CUDA events over a graph of 50 launches; one JSON line per shape."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
sys.path[:0] = [REPO, PKG]
import complexnn  # noqa: E402
from complexnn import _native  # noqa: E402



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


KERNELS = {0: "general", 1: "resident sub-filters (k_hamilton_tc)", 2: "streamed sub-filters (k_hamilton_tc2d)", 3: "small-K"}
args = [int(v) for v in sys.argv[1:]]
for i in range(0, len(args) - 4, 5):
    B, T, in_q, F, k = args[i:i + 5]
    np.random.seed(0)
    layer = complexnn.QuaternionConv1D(F, k, padding="same", activation="relu")
    x = torch.randn(B, T, 4 * in_q, device="cuda")
    layer(x)
    d = _native.make_conv_desc(1, B, (T,), in_q, F, (k,), (1,), (1,), "same", "channels_last", "relu")
    kern = _native.lib().qnn_conv_forward_kernel(ctypes.byref(d))
    us = timed(lambda: layer(x))
    qmac = B * T * k * in_q * F
    print(json.dumps({"shape": "x[%d,%d,4x%d] -> %d filters, k=%d" % (B, T, in_q, F, k), "kernel": KERNELS.get(kern, kern),
                      "us": us, "qMAC_per_s": qmac / us * 1e6, "TFLOPs": 32 * qmac / us / 1e6}))
