#!/usr/bin/env python
"""Timing of the general (CUDA-core) kernels on shapes the tensor-core kernel does not take (diagnostics)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
sys.path[:0] = [REPO, PKG]
import complexnn  # noqa: E402
from complexnn import _ops  # noqa: E402
from complexnn._layer import Variable  # noqa: E402


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


cases = [
    ("cfg5 slice conv2d NCHW B=8 64->128q 3x3 128x128", dict(x=(8, 256, 128, 128), k=(3, 3, 64, 512), F=128, cf=True)),
    ("cfg5 slice conv2d NHWC B=8", dict(x=(8, 128, 128, 256), k=(3, 3, 64, 512), F=128, cf=False)),
    ("timit first conv1d in_q=41 B=256 T=256", dict(x=(256, 256, 164), k=(3, 41, 256), F=64, cf=False)),
    ("decoda first conv1d in_q=1 B=325 T=250", dict(x=(325, 250, 4), k=(3, 1, 128), F=32, cf=False)),
    ("cfg2 on the general kernel", dict(x=(256, 256, 160), k=(3, 40, 256), F=64, cf=False)),
]
for name, c in cases:
    x = torch.randn(c["x"], device="cuda")
    kern = Variable((np.random.randn(*c["k"]) * 0.05).astype(np.float32))
    bias = Variable(np.zeros(4 * c["F"], np.float32))
    rank = len(c["k"]) - 2
    fmt = "channels_first" if c["cf"] else "channels_last"
    f = lambda: _ops.conv_forward(x, kern, bias, c["F"], c["k"][:rank], (1,) * rank, "same", fmt, (1,) * rank, "relu",
                                  math="fp32", algo="general")
    ms = timeit(f)
    y = f()
    P = int(np.prod(y.shape)) // (4 * c["F"])
    qmac = P * int(np.prod(c["k"][:rank])) * c["k"][rank] * c["F"]
    print("%-55s %9.3f ms  %7.2f TFLOP/s" % (name, ms, 32 * qmac / ms / 1e9))
    y2 = _ops.conv_forward(x, kern, bias, c["F"], c["k"][:rank], (1,) * rank, "same", fmt, (1,) * rank, "relu",
                           math="fp32", algo="general")
    dy = torch.randn_like(y2)
    g = lambda: _ops.conv_backward(x, y2, dy, kern, True, c["F"], c["k"][:rank], (1,) * rank, "same", fmt, (1,) * rank, "relu")
    ms = timeit(g, 2)
    print("%-55s %9.3f ms  %7.2f TFLOP/s (backward: dgrad+wgrad+bgrad)" % ("", ms, 64 * qmac / ms / 1e9))
