"""Sustained run of the channels_first tensor-core kernel with NVML clock / power sampling: python tools/tc2d_clocks.py [B] [seconds]"""
import os
import sys
import threading
import time

import numpy as np
import torch
import pynvml

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
sys.path[:0] = [REPO, PKG]
from complexnn import _ops  # noqa: E402
from complexnn._layer import Variable  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []
stop = False


def sampler():
    while not stop:
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.01)


x = torch.randn(B, 256, 128, 128, device="cuda")
kern = Variable((np.random.default_rng(0).normal(size=(3, 3, 64, 512)) / 48).astype(np.float32))
bias = Variable(np.zeros(512, np.float32))
args = (kern, bias, 128, (3, 3), (1, 1), "same", "channels_first", (1, 1), "relu")
_ops.conv_forward(x, *args, math="tf32", algo="tensor")
torch.cuda.synchronize()
th = threading.Thread(target=sampler)
th.start()
t0 = time.time()
n = 0
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
while time.time() - t0 < secs:
    for _ in range(5):
        _ops.conv_forward(x, *args, math="tf32", algo="tensor")
        n += 1
    torch.cuda.synchronize()
ev1.record()
torch.cuda.synchronize()
stop = True
th.join()
ms = ev0.elapsed_time(ev1) / n
qmac = B * 128 * 128 * 9 * 64 * 128
clk = np.array([s[1] for s in samples[len(samples) // 4:]])
pw = np.array([s[2] for s in samples[len(samples) // 4:]])
reasons = 0
for s in samples:
    reasons |= s[3]
print("B=%d: %.3f ms/call, %.1f TFLOP/s; SM clock median %d min %d max %d MHz; power median %.0f max %.0f W; throttle reasons 0x%x"
      % (B, ms, qmac * 32 / ms * 1e-9, np.median(clk), clk.min(), clk.max(), np.median(pw), pw.max(), reasons))
