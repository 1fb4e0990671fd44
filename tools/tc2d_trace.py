#!/usr/bin/env python
"""Per-CTA pipeline timeline of the channels_first tensor-core kernel (diagnostics): python tools/tc2d_trace.py [B]"""
import ctypes
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
sys.path[:0] = [REPO, PKG]
import complexnn  # noqa: E402
from complexnn import _native  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
np.random.seed(0)
layer = complexnn.QuaternionConv2D(128, (3, 3), padding="same", data_format="channels_first", activation="relu")
x = torch.randn(B, 256, 128, 128, device="cuda")
for _ in range(2):
    layer(x)
torch.cuda.synchronize()
buf = torch.zeros(148 * 256, dtype=torch.int64, device="cuda")
lib = _native.lib()
lib.qnn_debug_trace(ctypes.c_void_p(buf.data_ptr()), buf.numel() * 8)
layer(x)
torch.cuda.synchronize()
lib.qnn_debug_trace(None, 0)
t = buf.cpu().numpy().reshape(148, 256)
dur = t[:, 2] - t[:, 0]
print("CTA duration cycles: min %d median %d max %d" % (dur.min(), np.median(dur), dur.max()))
for cta in (0, 73):
    base = t[cta, 0]
    print("---- CTA %d: setup %d end %d" % (cta, t[cta, 1] - base, t[cta, 2] - base))
    for k in range(12):
        v = t[cta, 8 + 4 * k: 12 + 4 * k]
        if v[0]:
            print("  item %2d: mma start %8d  committed %8d  epi got acc %8d  stored %8d" % ((k,) + tuple(int(a - base) for a in v)))
    print("  second item, issuer per slot (a_full passed, b_full passed), relative to item start:")
    s0 = t[cta, 8 + 4]
    print("   ", " ".join("%d/%d" % (t[cta, 64 + 2 * i] - s0, t[cta, 65 + 2 * i] - s0) for i in range(48) if t[cta, 64 + 2 * i]))
    print("  second item, converter group 0 own stages (x_full, a_empty, converted, refilled):")
    print("   ", " ".join("%d/%d/%d/%d" % tuple(int(a - s0) for a in t[cta, 160 + 4 * k: 164 + 4 * k]) for k in range(24)
                          if t[cta, 160 + 4 * k]))
