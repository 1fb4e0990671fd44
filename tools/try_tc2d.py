"""Bring-up / timing script for the channels_first tensor-core kernel (run on the GPU box):
   python tools/try_tc2d.py [--time] [--big]"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
sys.path[:0] = [REPO, PKG]
from complexnn import _ops  # noqa: E402
from complexnn._layer import Variable  # noqa: E402
from oracle import qoracle as O  # noqa: E402

CASES = [
    # B, Q, F, spatial, k, d, pad, act, bias
    (1, 8, 32, (4, 128), (3, 3), (1, 1), "same", "linear", False),
    (2, 16, 64, (5, 132), (3, 5), (1, 1), "same", "relu", True),
    (1, 8, 128, (9, 64), (3, 5), (2, 1), "valid", "relu", True),
    (3, 24, 96, (300,), (3,), (2,), "same", "tanh", True),
    (2, 8, 32, (7, 40), (1, 1), (1, 1), "valid", "linear", True),
    (1, 64, 128, (12, 128), (3, 3), (1, 1), "same", "relu", True),
]


def run_case(c):
    B, Q, F, sp, k, d, pad, act, use_bias = c
    rng = np.random.default_rng(abs(hash(c)) % (2 ** 31))
    rank = len(sp)
    x = rng.normal(size=(B, 4 * Q) + sp).astype(np.float32)
    kern = (rng.normal(size=k + (Q, 4 * F)) / np.sqrt(4 * Q * np.prod(k))).astype(np.float32)
    bias = rng.normal(0, 0.1, size=4 * F).astype(np.float32) if use_bias else None
    ones = (1,) * rank
    y = _ops.conv_forward(torch.from_numpy(x).cuda(), Variable(kern), Variable(bias) if use_bias else None, F, k, ones,
                          pad, "channels_first", d, act, math="tf32", algo="tensor")
    torch.cuda.synchronize()
    ref = O.qconv_forward(x, kern, bias, F, ones, pad, "channels_first", d, act)
    bound = O.qconv_abs_bound(x, kern, F, ones, pad, "channels_first", d)
    yn = y.cpu().numpy()
    efro = float(np.linalg.norm(yn - ref) / (np.linalg.norm(ref) + 1e-30))
    ratio = float((np.abs(yn - ref) / (bound + 1e-3)).max())
    ok = efro <= 1e-3 and np.all(np.abs(yn - ref) <= 1e-3 * bound + 1e-6)
    print("%s  fro-rel %.3e  worst |d|/bound %.3e  %s" % (c, efro, ratio, "OK" if ok else "FAIL"), flush=True)
    return ok


def time_cfg5(B):
    x = torch.randn(B, 256, 128, 128, device="cuda")
    kern = Variable((np.random.default_rng(0).normal(size=(3, 3, 64, 512)) / 48).astype(np.float32))
    bias = Variable(np.zeros(512, np.float32))
    args = (kern, bias, 128, (3, 3), (1, 1), "same", "channels_first", (1, 1), "relu")
    for algo in ("tensor",) + (("general",) if B <= 8 else ()):
        y = _ops.conv_forward(x, *args, math="tf32" if algo == "tensor" else "fp32", algo=algo)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        n = 10 if algo == 'tensor' else 2
        ev[0].record()
        for _ in range(n):
            y = _ops.conv_forward(x, *args, math="tf32" if algo == "tensor" else "fp32", algo=algo)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / n
        qmac = B * 128 * 128 * 9 * 64 * 128
        print("cfg5 B=%d %s: %.3f ms  %.3e qMAC/s  %.1f TFLOP/s" % (B, algo, ms, qmac / ms * 1e3, qmac * 32 / ms * 1e-9),
              flush=True)


if __name__ == "__main__":
    ok = True
    if "--notest" not in sys.argv:
        for c in CASES:
            ok &= run_case(c)
    if "--time" in sys.argv:
        time_cfg5(8)
        if "--big" in sys.argv:
            time_cfg5(128)
    sys.exit(0 if ok else 1)
