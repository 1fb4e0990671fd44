"""Timing of the tensor-core kernel gradient alone (linear activation, no bias, no dx): python tools/try_wgrad.py"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
sys.path[:0] = [REPO, PKG]
from complexnn import _ops  # noqa: E402
from complexnn._layer import Variable  # noqa: E402

CASES = [("cfg2 conv in_q=40 F=64 k=3", 256, 256, 40, 64, 3), ("cfg3 conv in_q=64 F=64 k=3", 256, 256, 64, 64, 3),
         ("dense 64->64 (k=1)", 256, 256, 64, 64, 1), ("timit first layer in_q=41", 256, 256, 41, 64, 3)]
for name, B, T, in_q, F, k in CASES:
    x = torch.randn(B, T, 4 * in_q, device="cuda")
    dy = torch.randn(B, T, 4 * F, device="cuda")
    kern = Variable((np.random.default_rng(0).normal(size=(k, in_q, 4 * F)) * 0.05).astype(np.float32))
    args = (x, dy, dy, kern, False, F, (k,), (1,), "same", "channels_last", (1,), "linear")
    for algo in ("tensor",):
        _ops.conv_backward(*args, need_dx=False, math="tf32", algo=algo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            _ops.conv_backward(*args, need_dx=False, math="tf32", algo=algo)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        qm = B * T * k * in_q * F
        print("%-32s %s: %8.1f us  %.2e qMAC/s  %.0f TFLOP/s  (x + dz = %.0f MB -> %.2f TB/s)" % (
            name, algo, us, qm / us * 1e6, qm * 32 / us * 1e-6, (x.numel() + dy.numel()) * 4e-6,
            (x.numel() + dy.numel()) * 4 / us * 1e-6), flush=True)
