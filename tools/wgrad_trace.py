#!/usr/bin/env python
"""Per-CTA pipeline timeline of the tensor-core kernel-gradient kernel (diagnostics): python tools/wgrad_trace.py [in_q] [k]"""
import ctypes
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
sys.path[:0] = [REPO, PKG]
from complexnn import _native, _ops  # noqa: E402
from complexnn._layer import Variable  # noqa: E402

in_q = int(sys.argv[1]) if len(sys.argv) > 1 else 64
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, T, F = 256, 256, 64
x = torch.randn(B, T, 4 * in_q, device="cuda")
dy = torch.randn(B, T, 4 * F, device="cuda")
kern = Variable((np.random.default_rng(0).normal(size=(k, in_q, 4 * F)) * 0.05).astype(np.float32))
args = (x, dy, dy, kern, False, F, (k,), (1,), "same", "channels_last", (1,), "linear")
for _ in range(2):
    _ops.conv_backward(*args, need_dx=False, math="tf32", algo="tensor")
torch.cuda.synchronize()
buf = torch.zeros(148 * 256, dtype=torch.int64, device="cuda")
lib = _native.lib()
lib.qnn_debug_trace(ctypes.c_void_p(buf.data_ptr()), buf.numel() * 8)
_ops.conv_backward(*args, need_dx=False, math="tf32", algo="tensor")
torch.cuda.synchronize()
lib.qnn_debug_trace(None, 0)
t = buf.cpu().numpy().reshape(148, 256)
dur = t[:, 2] - t[:, 0]
print("CTA duration cycles: min %d median %d max %d" % (dur.min(), np.median(dur), dur.max()))
for cta in (0, 73):
    base = t[cta, 0]
    print("---- CTA %d: setup %d, accumulators complete %d, epilogue done %d, end %d" % (
        cta, t[cta, 1] - base, t[cta, 3] - base, t[cta, 4] - base, t[cta, 2] - base))
    print("  unit: issuer b_full / committed | packer b_empty / stored | converter x_full / done | producer x_empty")
    for u in range(24):
        v = t[cta, 8 + 8 * u: 8 + 8 * u + 7]
        if v[0]:
            r = [int(a - base) if a else -1 for a in v]
            print("  %2d: %7d %7d | %7d %7d | %7d %7d | %7d" % ((u,) + tuple(r)))
