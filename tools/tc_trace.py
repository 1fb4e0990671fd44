#!/usr/bin/env python
"""Per-CTA pipeline timeline of the tensor-core kernel (diagnostics).  Runs one forward with qnn_debug_trace enabled
and prints, for a few CTAs, the clock64() of each pipeline event relative to the CTA's start (in SM cycles).
  python tools/tc_trace.py [cfg2|dense|timit|few]"""
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

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
np.random.seed(0)
if which in ("cfg2", "few"):
    layer = complexnn.QuaternionConv1D(64, 3, padding="same", activation="relu")
    x = torch.randn(256 if which == "cfg2" else 2, 256, 160, device="cuda")   # "few": 4 tiles -> 4 CTAs, same weights
elif which == "timit":   # first layer of the cfg 3 stack: 41 TIMIT features per component (ragged channel count)
    layer = complexnn.QuaternionConv1D(64, 3, padding="same", activation="relu")
    x = torch.randn(256, 256, 164, device="cuda")
else:
    layer = complexnn.QuaternionDense(256, activation="relu")
    x = torch.randn(65536, 160, device="cuda")
for _ in range(3):
    layer(x)
torch.cuda.synchronize()
buf = torch.zeros(148 * 256, dtype=torch.int64, device="cuda")
lib = _native.lib()
lib.qnn_debug_trace(ctypes.c_void_p(buf.data_ptr()), buf.numel() * 8)
layer(x)
torch.cuda.synchronize()
lib.qnn_debug_trace(None, 0)
t = buf.cpu().numpy().reshape(148, 256)
names = {0: "start", 1: "setup done", 2: "packed", 3: "first TMA", 4: "TMA done", 5: "first x landed", 6: "w_ready seen",
         7: "first A slot", 58: "end"}
for k in range(5):
    for j, n in enumerate(["mma:acc_empty", "mma:committed", "epi:acc_full", "epi:tmem_free", "epi:stored"]):
        names[8 + 5 * k + j] = "tile%d %s" % (k, n)
t = t[t[:, 58] != 0]
g0 = t[:, 59].min()
print("globaltimer: first CTA start .. last CTA end = %.2f us; CTA start skew max %.2f us" % (
    (t[:, 60].max() - g0) / 1e3, (t[:, 59].max() - g0) / 1e3))
t = t[t[:, 58] != 0]
dur = t[:, 58] - t[:, 0]
print("CTA duration cycles: min %d median %d max %d" % (dur.min(), np.median(dur), dur.max()))
ncta = int((t[:, 58] != 0).sum())
print("CTAs that ran:", ncta)
for cta in ((0, 1, 73, 147) if ncta == 148 else (0, ncta - 1)):
    print("---- CTA %d (sm %d)" % (cta, t[cta, 61]))
    ev = sorted((int(t[cta, s] - t[cta, 0]), names[s]) for s in names if t[cta, s] != 0 or s == 0)
    for c, n in ev:
        print("  %8d  %s" % (c, n))

# detailed events of CTA 0's second tile
cta = 0
base = t[cta, 0]
print("==== CTA 0, second tile, detailed (cycles since CTA start)")
ev = []
for s_ in range(8):
    for j, n in enumerate(["conv x_full", "conv a_empty tap0", "conv a_empty tap1", "conv a_empty tap2", "conv a_empty tap3",
                           "conv wait::st done", "conv arrived"]):
        v = t[cta, 64 + 8 * s_ + j]
        if v:
            ev.append((int(v - base), "stage %d %s" % (s_, n)))
for sl in range(32):
    for j, n in enumerate(["issue a_full", "issue committed"]):
        v = t[cta, 128 + 2 * sl + j]
        if v:
            ev.append((int(v - base), "slot %d %s" % (sl, n)))
for s_ in range(16):
    for j, n in enumerate(["prod x_empty", "prod issued"]):
        v = t[cta, 192 + 2 * s_ + j]
        if v:
            ev.append((int(v - base), "stage %d %s" % (s_, n)))
for j, n in enumerate(["pack: raw sub-filters landed", "-", "pack: stored", "pack: fenced+arrived", "pack w0 it0", "pack w0 it1",
                       "pack w0 it2", "pack w0 it3", "pack w15 it0", "pack w15 it1", "pack w15 it2", "pack w15 it3"]):
    v = t[cta, 240 + j]
    if v:
        ev.append((int(v - base), n))
for c, n in sorted(ev):
    print("  %8d  %s" % (c, n))
