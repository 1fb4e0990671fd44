#!/usr/bin/env python
"""Headline benchmark: quaternion-MACs/s of the fused Hamilton conv forward (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload cfg2|dense|stack|train|cfg5] [--math tf32|3xtf32] [--no-secondary]

A step of the default workload = one QuaternionConv1D forward (64 filters, kernel 3, stride 1, `same`, bias, relu) over a
synthetic x[256, 256, 160] fp32 batch -- ONE launch of the fused tensor-core kernel through the layer API / C ABI.
Other workloads: `dense` (north-star QuaternionDense shape), `stack` (BASELINE configs[2]: 3 x QConv1D + 2 x QDense
forward), `train` (configs[3]: that stack forward + backward + NCCL bucket all-reduce), `cfg5` (configs[4]:
QuaternionConv2D channels_first).  N > 1 (under torchrun): every rank runs the same per-GPU batch on its own GPU (weak
scaling, batch-sharded data parallelism; the forward has no collective, `train` all-reduces one gradient bucket);
value = all ranks' qMACs / max-over-ranks time.

How the K steps are timed (any K): all K steps are captured into ONE CUDA graph (no eager remainder, no interpreter
between two ~30 us launches); right before the timed region the same graph is replayed once without synchronising, so
the GPU is busy when `e0` is recorded; then e0 | replay (exactly K steps) | e1 on the launching stream.  A second,
longer loop (>= 0.25 s of back-to-back replays, one CUDA-event pair per replay) gives the median the clocks are sampled
under (`sustained`).  `--impl reference` times the CPU restatement of the reference path (oracle/qoracle.py: slice ->
negate -> concatenate, then im2col + sgemm in NumPy or torch's CPU conv / matmul (oneDNN), whichever is faster on the
box; bias, relu; fp32) on the host cores on the SAME configuration.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
LOCAL_WORLD = int(os.environ.get("LOCAL_WORLD_SIZE", str(WORLD)))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


ALL_CORES = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))

# torchrun exports OMP_NUM_THREADS=1 to every worker; the CPU baseline / reference arm on rank 0 must see all host
# cores, and BLAS reads these variables when NumPy is first imported -- so set them before that import.
if RANK == 0:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(len(ALL_CORES))

# rank 0 prints ONE JSON line on stdout.  NCCL prints its "NCCL version ..." banner on stdout at NCCL_DEBUG=VERSION and
# WARN (the GPU boxes export VERSION) the first time ANY communicator is created in a process -- torch's or the
# library's own: drop the variable unless the user asked for real diagnostics, and (belt and braces) point file
# descriptor 1 of every rank at stderr for the whole run; the JSON line goes to the saved, real stdout.
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    del os.environ["NCCL_DEBUG"]
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "quaternion-MACs/sec (QConv1D+QDense fwd)"
WORKLOADS = {
    "cfg2": dict(kind="conv1d", B=256, T=256, in_q=40, F=64, k=3,
                 desc="QuaternionConv1D fwd x[256,256,4x40] 64 filters k=3 same relu (BASELINE configs[1])"),
    "dense": dict(kind="dense", B=65536, T=1, in_q=40, F=64, k=1,
                  desc="QuaternionDense fwd x[65536,160] -> 256 relu (north-star dense shape)"),
    "stack": dict(kind="stack", B=256, T=256,
                  desc="QCNN stack fwd: 3 x QConv1D(64,3,same,relu) + 2 x QDense(256,relu) on x[256,256,4x41] (BASELINE configs[2])"),
    "train": dict(kind="train", B=256, T=256,
                  desc="QCNN stack training step: fwd + bwd (dgrad, wgrad, dbias) + gradient-bucket all-reduce, "
                       "x[256,256,4x41] per GPU (BASELINE configs[3])"),
    # BASELINE configs[4]: T = H*W positions, k = 3*3 taps
    "cfg5": dict(kind="conv2d", B=128, T=128 * 128, H=128, W=128, in_q=64, F=128, k=9,
                 desc="QuaternionConv2D fwd x[128,4x64,128,128] channels_first 128 filters k=3x3 same relu (BASELINE configs[4])"),
}
STACK_IN_Q = [41, 64, 64, 64, 64]
STACK_TAPS = [3, 3, 3, 1, 1]


def qmacs(w, batch=None):
    B = batch if batch is not None else w["B"]
    if w["kind"] in ("stack", "train"):
        fwd = sum(B * w["T"] * t * q * 64 for t, q in zip(STACK_TAPS, STACK_IN_Q))
        if w["kind"] == "stack":
            return fwd
        return 3 * fwd - B * w["T"] * STACK_TAPS[0] * STACK_IN_Q[0] * 64      # the first layer has no data gradient
    return B * w["T"] * w["k"] * w["in_q"] * w["F"]


def alg_bytes(w):
    """Algorithmic HBM bytes per step: every layer reads its input once, writes its output once, reads the stored
    (un-expanded) kernel and bias; the training step additionally reads dy / y / x and writes dx per layer."""
    if w["kind"] in ("stack", "train"):
        rows = w["B"] * w["T"]
        acts = [164, 256, 256, 256, 256, 256]
        fwd = sum(4 * rows * (acts[i] + acts[i + 1]) for i in range(5))
        wts = sum(4 * (t * q * 256 + 256) for t, q in zip(STACK_TAPS, STACK_IN_Q))
        if w["kind"] == "stack":
            return fwd + wts
        bwd = sum(4 * rows * (2 * acts[i + 1] + acts[i] + (acts[i] if i > 0 else 0)) for i in range(5))
        return fwd + bwd + 3 * wts
    return 4 * (w["B"] * w["T"] * 4 * w["in_q"] + w["B"] * w["T"] * 4 * w["F"] + w["k"] * w["in_q"] * 4 * w["F"] + 4 * w["F"])


def peaks():
    try:
        m = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        return dict(hbm_gbs=float(m["hbm_gbs"]), bf16=float(m["bf16_tflops"]),
                    bf16_sustained=float(m.get("bf16_tflops_sustained", m["bf16_tflops"])), src="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons through NVML (nvidia-smi as a fallback) at <= 50 Hz from BEFORE the
    warm-up until after the timed loops; `summary(t0, t1)` reports the samples that fall inside a window."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period, self.rows, self._stop_evt = index, period, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample(self):
        t = time.perf_counter()
        if self.nvml is not None:
            n = self.nvml
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                try:
                    pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:
                    pw = None
                flags = [bool(r & n.nvmlClocksThrottleReasonHwSlowdown), bool(r & n.nvmlClocksThrottleReasonHwThermalSlowdown),
                         bool(r & n.nvmlClocksThrottleReasonSwThermalSlowdown), bool(r & n.nvmlClocksThrottleReasonSwPowerCap)]
                self.rows.append((t, float(sm), float(mx), pw, flags))
                return
            except Exception:
                self.nvml = None
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            c = [v.strip() for v in out.strip().split(",")]
            self.rows.append((t, float(c[0]), float(c[1]), float(c[2]), [v.lower().startswith("active") for v in c[3:7]]))
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            self.sample()
            self._stop_evt.wait(self.period if self.nvml is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)

    def summary(self, t0, t1):
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-1:]
        reasons = sorted({n for r in rows for n, f in zip(self.NAMES, r[4]) if f})
        pw = [r[3] for r in rows if r[3] is not None]
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None,
                "sm_max_mhz": max(r[2] for r in rows) if rows else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(rows),
                "window_s": round(t1 - t0, 4), "source": "nvml" if self.nvml is not None else "nvidia-smi",
                "period_s": self.period}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm (the oracle's fp32 port of the reference path; test infrastructure used here as the timed baseline only)
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_step(w, sample_batch, rng):
    """The CPU restatements of the reference path on `sample_batch` units: {name: closure} and the qMACs of one pass.
    "numpy" is the literal port (slice -> negate -> concatenate -> im2col -> sgemm); "torch-cpu" hands the real
    conv / matmul to torch's CPU kernels (oneDNN / MKL), the closest thing here to TensorFlow's CPU backend."""
    from oracle import qoracle as O
    if w["kind"] == "conv1d":
        x = rng.normal(size=(sample_batch, w["T"], 4 * w["in_q"])).astype(np.float32)
        kern = (rng.normal(size=(w["k"], w["in_q"], 4 * w["F"])) * 0.05).astype(np.float32)
        bias = rng.normal(0, 0.1, 4 * w["F"]).astype(np.float32)
        return {"numpy": lambda: O.qconv1d_forward_f32(x, kern, bias, w["F"], "same", True),
                "torch-cpu": lambda: O.qconv1d_forward_torch_cpu(x, kern, bias, w["F"], "same", True)}, qmacs(w, sample_batch)
    if w["kind"] == "conv2d":
        x = rng.normal(size=(sample_batch, 4 * w["in_q"], w["H"], w["W"])).astype(np.float32)
        kern = (rng.normal(size=(3, 3, w["in_q"], 4 * w["F"])) * 0.05).astype(np.float32)
        bias = rng.normal(0, 0.1, 4 * w["F"]).astype(np.float32)
        return {"numpy": lambda: O.qconv2d_forward_f32(x, kern, bias, w["F"], True),
                "torch-cpu": lambda: O.qconv2d_forward_torch_cpu(x, kern, bias, w["F"], True)}, qmacs(w, sample_batch)
    if w["kind"] in ("stack", "train"):
        # forward of the stack (the CPU arm of `train` is the forward too: the reference's backward is TF autodiff)
        x = rng.normal(size=(sample_batch, w["T"], 164)).astype(np.float32)
        ks = [(rng.normal(size=(3, q, 256)) * 0.05).astype(np.float32) for q in STACK_IN_Q[:3]]
        kd = [(rng.normal(size=(64, 256)) * 0.05).astype(np.float32) for _ in range(2)]
        bs = [rng.normal(0, 0.1, 256).astype(np.float32) for _ in range(5)]

        def run(conv, dense):
            h = x
            for i in range(3):
                h = conv(h, ks[i], bs[i], 64, "same", True)
            h = h.reshape(-1, 256)
            for i in range(2):
                h = dense(h, kd[i], bs[3 + i], 256, True)
            return h
        return {"numpy": lambda: run(O.qconv1d_forward_f32, O.qdense_forward_f32),
                "torch-cpu": lambda: run(O.qconv1d_forward_torch_cpu, O.qdense_forward_torch_cpu)}, \
            qmacs(dict(w, kind="stack"), sample_batch)
    x = rng.normal(size=(sample_batch, 4 * w["in_q"])).astype(np.float32)
    kern = (rng.normal(size=(w["in_q"], 4 * w["F"])) * 0.05).astype(np.float32)
    bias = rng.normal(0, 0.1, 4 * w["F"]).astype(np.float32)
    return {"numpy": lambda: O.qdense_forward_f32(x, kern, bias, 4 * w["F"], True),
            "torch-cpu": lambda: O.qdense_forward_torch_cpu(x, kern, bias, 4 * w["F"], True)}, qmacs(w, sample_batch)


def fastest_cpu_step(cands):
    """Times every candidate briefly (one warm-up + >= 0.5 s) and returns (name, closure, {name: seconds per pass}):
    the CPU arm is the FASTEST restatement available on the box, not the most convenient one."""
    per = {}
    for name, fn in cands.items():
        try:
            fn()
            t0, n = time.perf_counter(), 0
            while n < 2 or time.perf_counter() - t0 < 0.5:
                fn()
                n += 1
            per[name] = (time.perf_counter() - t0) / n
        except Exception:          # e.g. torch without CPU conv support: keep the other candidate
            continue
    best = min(per, key=per.get)
    return best, cands[best], per


def cpu_sample(w):
    """Units of the workload one CPU-arm step processes.  cfg 2, the dense shape and the stack run WHOLE (the same
    configuration as the GPU arm: ~10-60 ms per step on the box's cores); only cfg 5 (5 TFLOP, minutes per step on a
    CPU) is a bounded sample."""
    return {"conv1d": w["B"], "conv2d": 1, "dense": w["B"], "stack": w["B"], "train": w["B"]}[w["kind"]]


def unit_name(w):
    return {"conv1d": "sequences", "conv2d": "images", "dense": "rows", "stack": "sequences", "train": "sequences"}[w["kind"]]


class all_host_threads(object):
    """The CPU arm uses every host core for its BLAS calls, also under torchrun (which exports OMP_NUM_THREADS=1)."""

    def __enter__(self):
        try:
            os.sched_setaffinity(0, ALL_CORES)
        except Exception:
            pass
        try:
            import torch
            torch.set_num_threads(len(ALL_CORES))
        except Exception:
            pass
        try:
            from threadpoolctl import threadpool_limits
            self._ctx = threadpool_limits(limits=len(ALL_CORES))
            self._ctx.__enter__()
        except Exception:
            self._ctx = None
        return self

    def __exit__(self, *a):
        if self._ctx is not None:
            self._ctx.__exit__(*a)


def run_reference(args, w):
    """The reference arm: CPU restatement of complexnn/conv.py:288-345 / dense.py:126-164 on the box's host cores."""
    if RANK != 0:
        return 0
    rng = np.random.default_rng(0)
    sample = cpu_sample(w)
    cands, q = cpu_reference_step(w, sample, rng)
    with all_host_threads():
        best, step, per = fastest_cpu_step(cands)
        for _ in range(max(args.warmup, 1)):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
    val = q / dt
    whole = sample == w["B"]
    sample_desc = "%s per step%s; fastest of %s -> %s (expansion + conv/matmul + bias + relu, fp32)" % (
        ("the whole workload (%d %s)" % (w["B"], unit_name(w))) if whole else "%d of %d %s" % (sample, w["B"], unit_name(w)),
        "" if w["kind"] != "train" else " (forward of the stack: the reference's backward is TF autodiff, not restated on CPU)",
        {k: "%.3g qMAC/s" % (q / v) for k, v in per.items()}, best)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "qMAC/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "sample": sample_desc, "same_config": whole},
            "cpu_baseline": {"value": val, "unit": "qMAC/s", "cores": len(ALL_CORES), "kind": "port", "sample": sample_desc},
            "e2e": {"value": val, "unit": "qMAC/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def contract_stats(y, ref):
    """max|d|/max|ref| and the number of elements violating allclose(rtol=1e-3, atol=1e-3*rms(ref))."""
    y = np.asarray(y, np.float64)
    ref = np.asarray(ref, np.float64)
    d = np.abs(y - ref)
    rms = float(np.sqrt(np.mean(ref * ref)))
    return float(d.max() / np.abs(ref).max()), int((d > 1e-3 * rms + 1e-3 * np.abs(ref)).sum()), int(ref.size)


class Workload(object):
    """One benchmark workload on the current GPU: `step(i)` enqueues one step on input set i % n_sets."""

    def __init__(self, name, w, math, rank, world):
        import torch
        import complexnn
        self.name, self.w, self.math = name, w, math
        self.torch = torch
        rng = np.random.default_rng(1234 + rank)
        np.random.seed(0)
        kind = w["kind"]
        os.environ["QNN_MATH"] = math
        self.bucket = None
        if kind == "conv1d":
            self.layers = [complexnn.QuaternionConv1D(w["F"], w["k"], padding="same", activation="relu")]
            in_shape = (w["B"], w["T"], 4 * w["in_q"])
        elif kind == "conv2d":
            self.layers = [complexnn.QuaternionConv2D(w["F"], (3, 3), padding="same", data_format="channels_first",
                                                      activation="relu")]
            in_shape = (w["B"], 4 * w["in_q"], w["H"], w["W"])
        elif kind == "dense":
            self.layers = [complexnn.QuaternionDense(4 * w["F"], activation="relu")]
            in_shape = (w["B"], 4 * w["in_q"])
        else:
            self.layers = [complexnn.QuaternionConv1D(64, 3, padding="same", activation="relu") for _ in range(3)] + \
                          [complexnn.QuaternionDense(256, activation="relu") for _ in range(2)]
            in_shape = (w["B"], w["T"], 164)
        self.in_shape = in_shape
        # inputs resident in HBM; rotate buffer sets so no step finds its input in the 126 MB L2
        self.n_sets = 3 if alg_bytes(w) < 1e9 and kind != "train" else 1
        self.xs = [torch.randn(in_shape, device="cuda", dtype=torch.float32) for _ in range(self.n_sets)]
        # build + non-zero biases (the epilogue does real work)
        self.forward(self.xs[0])
        for layer in self.layers:
            ws = layer.get_weights()
            ws[-1] = rng.normal(0, 0.1, ws[-1].shape).astype(np.float32)
            layer.set_weights(ws)
        if kind == "train":
            from complexnn.dataparallel import GradBucket
            self.bucket = GradBucket(self.layers, device="cuda")
            rows = w["B"] * w["T"]
            self.dy = torch.randn(rows, 256, device="cuda") / rows
        self.world = world
        self.y = None

    def forward(self, x, keep=None):
        h = x
        B, T = self.w["B"], self.w["T"]
        for i, layer in enumerate(self.layers):
            if self.w["kind"] in ("stack", "train") and i == 3:
                h = h.reshape(B * T, 256)
            if keep is not None:
                keep.append(h)
            h = layer(h)
        if keep is not None:
            keep.append(h)
        return h

    def step(self, i):
        x = self.xs[i % self.n_sets]
        if self.w["kind"] != "train":
            self.y = self.forward(x)
            return
        acts = []
        self.y = self.forward(x, keep=acts)
        g = self.dy
        B, T = self.w["B"], self.w["T"]
        for li in (4, 3, 2, 1, 0):
            layer = self.layers[li]
            dk, db = self.bucket.views(layer)
            if li == 2:
                g = g.view(B, T, 256)
            g, _, _ = layer.backward(acts[li], acts[li + 1] if li != 2 else acts[3].view(B, T, 256), g,
                                     need_input_grad=li > 0, grad_kernel_out=dk, grad_bias_out=db)

    def exchange(self):
        """The one collective of the training step: in-place sum of the flat gradient bucket over ranks (NCCL)."""
        if self.bucket is not None and self.world > 1:
            from complexnn.dataparallel import allreduce_
            allreduce_(self.bucket, average_over=self.world)

    def parity(self):
        """Parity gate on the very tensors that are timed, against the oracle (SURVEY 8d metric)."""
        from oracle import qoracle as O
        w, kind = self.w, self.w["kind"]
        torch = self.torch
        if kind in ("stack", "train"):
            x = self.xs[0][:2]
            save = self.w
            self.w = dict(w, B=2)
            try:
                got = self.forward(x.contiguous()).cpu().numpy()
            finally:
                self.w = save
            ref = x.cpu().numpy()
            for i, layer in enumerate(self.layers):
                k, b = layer.get_weights()
                if i < 3:
                    ref = O.qconv_forward(ref, k, b, 64, 1, "same", "channels_last", 1, "relu")
                else:
                    ref = O.qdense_forward(ref.reshape(-1, 256), k, b, 256, "relu")
        else:
            sl = {"conv1d": slice(0, 4), "conv2d": slice(0, 1), "dense": slice(0, 1024)}[kind]
            xh = self.xs[0][sl]
            if kind == "conv2d":
                xh = xh[:, :, :16]          # 16 image rows of one sample: seconds, not minutes, for the fp64 oracle
            xh = xh.contiguous()
            got = self.layers[0](xh).cpu().numpy()
            k, b = self.layers[0].get_weights()
            if kind == "conv1d":
                ref = O.qconv_forward(xh.cpu().numpy(), k, b, w["F"], 1, "same", "channels_last", 1, "relu")
            elif kind == "conv2d":
                ref = O.qconv_forward(xh.cpu().numpy(), k, b, w["F"], (1, 1), "same", "channels_first", (1, 1), "relu")
            else:
                ref = O.qdense_forward(xh.cpu().numpy(), k, b, 4 * w["F"], "relu")
        max_rel, viol, n = contract_stats(got, ref)
        return {"mode": self.math, "max_rel": max_rel, "allclose_violations": viol, "checked": n,
                "criterion": "max|d|/max|ref| <= 1e-3 and allclose(rtol=1e-3, atol=1e-3*rms(ref)) vs the fp64-accumulated oracle"}


def time_workload(wl, steps, warmup, dist, min_sustain_s=0.25, sampler=None):
    """Times exactly `steps` steps of `wl` (CUDA events on the launching stream, GPU kept busy by an untimed pre-roll),
    then a sustained loop; returns a dict of timings (ms per step, max over ranks)."""
    import torch
    from complexnn import _native
    world = wl.world
    # Training across ranks: the NCCL all-reduce of the gradient bucket is captured INTO the graph with the step's kernels,
    # in "thread_local" capture mode (under torch's default "global" mode the capture hung in round 1: NCCL's own threads
    # make CUDA calls that mode forbids while any thread captures).  QNN_BENCH_NCCL_IN_GRAPH=0: graph of one step's
    # kernels, all-reduce launched after each replay.
    nccl_in_graph = os.environ.get("QNN_BENCH_NCCL_IN_GRAPH", "1") == "1"
    train_multi = wl.bucket is not None and world > 1 and not nccl_in_graph
    for i in range(max(warmup, 3)):
        wl.step(i)
        wl.exchange()
    torch.cuda.synchronize()
    # ---- capture: all `steps` steps in one graph (forward-only workloads), or ONE step (training across ranks: the
    # all-reduce stays outside the graph -- capturing NCCL of the library's own communicator hung in round 1)
    per_graph = 1 if train_multi else steps
    cap = torch.cuda.Stream()
    cap.wait_stream(torch.cuda.current_stream())
    l0 = _native.launch_count()
    with torch.cuda.stream(cap):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap, capture_error_mode="thread_local" if nccl_in_graph else "global"):
            for i in range(per_graph):
                wl.step(i)
                if nccl_in_graph:
                    wl.exchange()
    launches_per_step = (_native.launch_count() - l0) / float(per_graph)
    torch.cuda.current_stream().wait_stream(cap)

    def run_steps():
        if train_multi:
            for _ in range(steps):
                graph.replay()
                wl.exchange()
        else:
            graph.replay()

    run_steps()                     # first replay (uploads the graph)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_win0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pre = max(1, -(-20 // steps))   # >= 20 untimed steps queued right before e0, no synchronisation in between
    for _ in range(pre):
        run_steps()
    e0.record()
    run_steps()                     # EXACTLY `steps` steps
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # ---- sustained loop: back-to-back replays for >= min_sustain_s, one event pair per replay -> median
    reps = int(min(2000, max(11, np.ceil(min_sustain_s * 1e3 / max(ms * steps, 1e-3)))))
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for r in range(reps):
        run_steps()
        evs[r + 1].record()
    torch.cuda.synchronize()
    t_win1 = time.perf_counter()
    per = np.array([evs[r].elapsed_time(evs[r + 1]) / steps for r in range(reps)])
    med = float(np.median(per))
    if world > 1:
        t = torch.tensor([ms, med], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, med = float(t[0].item()), float(t[1].item())
        dist.barrier()
    out = {"ms_per_step": ms, "sustained_ms_per_step_median": med, "sustained_replays": reps,
           "sustained_seconds": float(per.sum() * steps / 1e3), "launches_per_step": launches_per_step,
           "window": (t_win0, t_win1),
           "launch": ("CUDA graph of one step replayed %d times, all-reduce launched after each replay" % steps) if train_multi
           else "ONE CUDA graph holding all %d steps%s (inputs rotate over %d sets); %d untimed steps queued right before "
                "the timed replay" % (steps, " incl. the NCCL all-reduce of every step" if wl.bucket is not None and world > 1
                                      else "", wl.n_sets, pre * steps)}
    return out


def measure_tf32_peak(seconds=1.5):
    """cuBLAS TF32 GEMM throughput on this GPU, now: burst (best of 10 single 8192^3 matmuls) and sustained (back-to-back
    for `seconds`).  This is the tensor-core denominator of the roofline for a kind::tf32 kernel."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device="cuda")
        b = torch.randn(n, n, device="cuda")
        c = torch.empty(n, n, device="cuda")
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        fl = 2.0 * n ** 3
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        return {"burst_tflops": fl / (best * 1e-3) / 1e12, "sustained_tflops": fl * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12,
                "how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS): best of 10 (burst), %d back to back (sustained)" % reps}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b, c
        torch.cuda.empty_cache()


def roofline_block(w, ms, pk, tf32, math, kernel):
    flops = 32.0 * qmacs(w) * (3 if math == "3xtf32" else 1)      # 3xTF32 issues three MMAs per block
    alg_flops = 32.0 * qmacs(w)
    t_s = ms * 1e-3
    long_step = ms > 2.0      # a multi-millisecond tensor-bound step runs at the power cap: sustained peak applies
    peak_tf = tf32["sustained_tflops" if long_step else "burst_tflops"] if tf32 else pk["bf16_sustained" if long_step else "bf16"] / 2.0
    t_tc = flops / (peak_tf * 1e12)
    t_hbm = alg_bytes(w) / (pk["hbm_gbs"] * 1e9)
    bound = "tensor" if t_tc >= t_hbm else "hbm"
    ach_tf = alg_flops / t_s / 1e12
    ach_gbs = alg_bytes(w) / t_s / 1e9
    blk = {"bound": bound,
           "achieved": ach_tf if bound == "tensor" else ach_gbs,
           "peak": peak_tf / (3 if math == "3xtf32" else 1) if bound == "tensor" else pk["hbm_gbs"],
           "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
           "frac": max(t_tc, t_hbm) / t_s, "traffic": None, "kernel": kernel,
           "peak_source": ("cuBLAS TF32 GEMM measured in this run (%s)%s" % ("sustained: multi-ms step at the power cap" if long_step else "burst",
                           "; 3xTF32 = three MMAs per algorithmic block, peak / 3" if math == "3xtf32" else "")) if tf32 else
                          "tf32 dense = 1/2 x bf16, " + pk["src"],
           "t_tensor_us": t_tc * 1e6, "t_hbm_us": t_hbm * 1e6, "t_measured_us": t_s * 1e6,
           "tensor_frac": t_tc / t_s, "hbm_frac": t_hbm / t_s,
           "tensor_achieved_tflops": ach_tf, "tensor_peak_tflops": peak_tf,
           "half_bf16_peak_tflops": pk["bf16"] / 2.0, "tensor_frac_of_half_bf16": (alg_flops / (pk["bf16"] / 2.0 * 1e12)) / t_s,
           "hbm_achieved_gbs": ach_gbs, "hbm_peak_gbs": pk["hbm_gbs"], "hbm_peak_source": pk["src"],
           "flops_per_launch": alg_flops, "algorithmic_bytes_per_launch": alg_bytes(w)}
    return blk


def pin_rank_cores():
    """Under torchrun give every rank its own slice of the host cores (launch threads of 8 ranks on 16-32 cores
    otherwise migrate and collide); rank 0's CPU-baseline leg widens its mask again (all_host_threads)."""
    if WORLD <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    n = max(1, len(ALL_CORES) // max(LOCAL_WORLD, 1))
    mine = ALL_CORES[LOCAL * n:(LOCAL + 1) * n] or ALL_CORES
    try:
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


def run_ours(args, w):
    import torch
    import torch.distributed as dist

    world, rank, local = WORLD, RANK, LOCAL
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    cores = pin_rank_cores()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import importlib.util
    spec = importlib.util.spec_from_file_location("qnn_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if rank == 0:
        mod.build()
    if world > 1:
        dist.barrier()
    from complexnn import _native

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    comm = False
    if w["kind"] == "train" and world > 1:
        from complexnn.dataparallel import init_comm
        init_comm(rank, world)
        comm = True

    wl = Workload(args.workload, w, args.math, rank, world)
    parity = wl.parity() if rank == 0 else None
    # a fast wrong kernel is not a result: one layer must meet max|d|/max|ref| <= 1e-3; through the 5-layer relu stack TF32
    # rounding compounds (and flips a few relu masks), there the gate is 5e-3
    gate = 1e-3 if w["kind"] in ("conv1d", "conv2d", "dense") else 5e-3
    if parity is not None and not parity["max_rel"] <= gate:
        raise SystemExit("parity check failed before timing: max-rel error %.3e" % parity["max_rel"])
    tm = time_workload(wl, args.steps, args.warmup, dist, sampler=sampler)
    ms = tm["ms_per_step"]
    clocks = sampler.summary(*tm["window"]) if sampler else None

    # ---- secondary workloads in the same run (driver-visible): the QDense half of the metric, the cfg 3 stack, the
    # training step, the parity-safe 3xTF32 mode of the headline shape
    secondary = {}
    if args.workload == "cfg2" and not args.no_secondary:
        for name, math in (("cfg2_3xtf32", "3xtf32"), ("dense", args.math), ("stack", args.math), ("train", args.math),
                           ("cfg5", args.math)):
            wname = "cfg2" if name.startswith("cfg2") else name
            w2 = WORKLOADS[wname]
            try:
                if wname == "train" and world > 1 and not comm:
                    from complexnn.dataparallel import init_comm
                    init_comm(rank, world)
                    comm = True
                wl2 = Workload(wname, w2, math, rank, world)
                par2 = wl2.parity() if rank == 0 else None
                t2 = time_workload(wl2, 5 if wname == "cfg5" else max(10, min(args.steps, 50)), 3, dist, min_sustain_s=0.1)
                if rank == 0:
                    secondary[name] = {"workload": w2["desc"], "math": math, "ms_per_step": t2["ms_per_step"],
                                       "sustained_ms_per_step_median": t2["sustained_ms_per_step_median"],
                                       "value": world * qmacs(w2) / (t2["ms_per_step"] * 1e-3), "unit": "qMAC/s",
                                       "launches_per_step": t2["launches_per_step"], "parity": par2, "launch": t2["launch"],
                                       "_w": wname, "_math": math}
                del wl2
                torch.cuda.empty_cache()
            except Exception as exc:          # a secondary block must never cost the headline line
                if rank == 0:
                    secondary[name] = {"error": repr(exc)[:300]}
    os.environ["QNN_MATH"] = args.math

    # ---- end to end through the C ABI's host-buffer entry point (what the layer calls for NumPy inputs):
    # pinned host x -> H2D -> kernel -> D2H -> pinned host y, every step, synchronous return
    e2e = None
    if w["kind"] in ("conv1d", "conv2d", "dense"):
        import ctypes
        e2e_steps = max(3, min(args.steps, 20)) if alg_bytes(w) < 1e9 else 3
        layer = wl.layers[0]
        ws = layer.get_weights()
        xh_t = torch.empty(wl.in_shape, dtype=torch.float32).pin_memory()
        xh_t.copy_(wl.xs[0])
        y_dev = layer(wl.xs[0])
        yh_t = torch.empty(tuple(y_dev.shape), dtype=torch.float32).pin_memory()
        lib = _native.lib()
        hp = lambda a: ctypes.c_void_p(a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data)
        m_id = _native.MATH[args.math]
        if w["kind"] == "conv1d":
            desc = _native.make_conv_desc(1, w["B"], (w["T"],), w["in_q"], w["F"], (w["k"],), (1,), (1,), "same",
                                          "channels_last", "relu", math=args.math)
            host_call = lambda: _native.check(lib.qnn_conv_forward_host(ctypes.byref(desc), hp(xh_t), hp(ws[0]), hp(ws[1]),
                                                                        hp(yh_t), None))
        elif w["kind"] == "conv2d":
            desc = _native.make_conv_desc(2, w["B"], (w["H"], w["W"]), w["in_q"], w["F"], (3, 3), (1, 1), (1, 1), "same",
                                          "channels_first", "relu", math=args.math)
            host_call = lambda: _native.check(lib.qnn_conv_forward_host(ctypes.byref(desc), hp(xh_t), hp(ws[0]), hp(ws[1]),
                                                                        hp(yh_t), None))
        else:
            host_call = lambda: _native.check(lib.qnn_dense_forward_host(w["B"], w["in_q"], w["F"], hp(xh_t), hp(ws[0]),
                                                                         hp(ws[1]), 1, m_id, 0, hp(yh_t), None))
        host_call()                                         # warm (uploads + packs the weights once; they stay resident)
        torch.cuda.synchronize()
        if rank == 0:
            e2e_err = float((yh_t[:2] - y_dev[:2].cpu()).abs().max())
            assert e2e_err == 0.0, "host-buffer path differs from the device path"
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_call()                                     # returns after the result landed on the host
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        # the copy floor of this box at this N: the same bytes, pinned, H2D and D2H on two streams at once, no kernel --
        # every rank at the same time (the ranks share the host's PCIe root complexes and memory controllers)
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        x_dev, y_dev2 = torch.empty_like(wl.xs[0]), torch.empty(tuple(y_dev.shape), device="cuda")
        def copies():
            with torch.cuda.stream(s_in):
                x_dev.copy_(xh_t, non_blocking=True)
            with torch.cuda.stream(s_out):
                yh_t.copy_(y_dev2, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()
        copies()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            copies()
        floor_s = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([floor_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            floor_s = float(t.item())
        del x_dev, y_dev2
        e2e = {"value": world * qmacs(w) / e2e_s, "unit": "qMAC/s",
               "h2d_bytes_per_step": int(xh_t.numel() * 4), "d2h_bytes_per_step": int(yh_t.numel() * 4),
               "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "copy_floor_ms": floor_s * 1e3, "frac_of_copy_floor": floor_s / e2e_s,
               "copy_floor": "the same H2D + D2H bytes from / to pinned memory on two streams at once, no kernel, all %d "
                             "ranks simultaneously (max over ranks): what the host side of this box allows at this N" % world,
               "weights": "kernel + bias stay resident on the device between calls (keyed by host pointer + content hash)"}
    else:
        # the stack / training step through the public layer API with HOST inputs: pinned x -> H2D -> step -> D2H of the result
        e2e_steps = max(3, min(args.steps, 10))
        xh_t = torch.empty(wl.in_shape, dtype=torch.float32).pin_memory()
        xh_t.copy_(wl.xs[0])
        out_host = torch.empty(tuple(wl.y.shape), dtype=torch.float32).pin_memory()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            wl.xs[0].copy_(xh_t, non_blocking=True)
            wl.step(0)
            wl.exchange()
            out_host.copy_(wl.y, non_blocking=True)
            torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": world * qmacs(w) / e2e_s, "unit": "qMAC/s", "h2d_bytes_per_step": int(xh_t.numel() * 4),
               "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps}

    if comm:
        from complexnn.dataparallel import destroy_comm
        destroy_comm()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- rank 0: peaks, roofline, CPU baseline (the other ranks wait at the barrier above)
    pk = peaks()
    tf32 = None
    try:
        tf32 = measure_tf32_peak()
    except Exception:
        tf32 = None
    kernel = {"conv2d": "k_hamilton_tc2d", "train": "k_hamilton_tc (forward + data gradient) + k_hamilton_wgrad_tc"}.get(
        w["kind"], "k_hamilton_tc")
    roof = roofline_block(w, ms, pk, tf32, args.math, kernel)
    try:
        roof["traffic"] = json.load(open(os.path.join(REPO, "profiles", "traffic.json"))).get(args.workload)
    except Exception:
        pass
    for name, blk in secondary.items():
        if "error" not in blk:
            wn = blk.pop("_w")
            r2 = roofline_block(WORKLOADS[wn], blk["ms_per_step"], pk, tf32, blk.pop("_math"),
                                "k_hamilton_tc2d" if wn == "cfg5" else "k_hamilton_tc")
            blk["roofline"] = {k: r2[k] for k in ("bound", "achieved", "peak", "unit", "frac", "tensor_frac", "hbm_frac")}
    cpu_block = None
    if world == 1:      # the CPU baseline is timed at N = 1 only (at N > 1 the other ranks' host threads share the cores)
        cands, cpu_q = cpu_reference_step(w, cpu_sample(w), np.random.default_rng(0))
        with all_host_threads():
            cpu_best, cpu_step, cpu_per = fastest_cpu_step(cands)
            t0 = time.perf_counter()
            reps = 0
            while reps < 3 or time.perf_counter() - t0 < 10.0:
                cpu_step()
                reps += 1
                if reps >= 500:
                    break
            cpu_val = cpu_q * reps / (time.perf_counter() - t0)
        whole = cpu_sample(w) == w["B"]
        cpu_block = {"value": cpu_val, "unit": "qMAC/s", "cores": len(ALL_CORES), "kind": "port",
                     "sample": "%d x (%s), fastest of %s -> %s (expansion + conv/matmul + bias + relu, fp32)" % (
                         reps, ("the whole workload, %d %s" % (w["B"], unit_name(w))) if whole else
                         "%d of %d %s" % (cpu_sample(w), w["B"], unit_name(w)),
                         {k: "%.3g qMAC/s" % (cpu_q / v) for k, v in cpu_per.items()}, cpu_best)}

    q = qmacs(w)
    t_s = ms * 1e-3
    line = {
        "metric": METRIC, "value": world * q / t_s, "unit": "qMAC/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "tf32" if args.math == "tf32" else "3xtf32", "data": "synthetic",
        "config": {"workload": w["desc"], "per_gpu_batch": w["B"], "global_batch": w["B"] * world,
                   "parallelism": ("dp%d (batch shards, no collective in forward)" % world) if w["kind"] != "train" else
                                  ("dp%d (batch shards; one NCCL all-reduce of the %d-float gradient bucket per step)" % (
                                      world, wl.bucket.numel())),
                   "l2": "rotating %d input sets (%.0f MB > 126 MB L2), output rewritten each step" % (
                       wl.n_sets, wl.n_sets * alg_bytes(w) / 1e6) if wl.n_sets > 1 else
                         "working set %.0f MB per step > 126 MB L2" % (alg_bytes(w) / 1e6),
                   "math": {"tf32": "tf32 operands (round-to-nearest), fp32 accumulate, fp32 I/O",
                            "3xtf32": "3xTF32: hi/lo operand split, three MMAs per block, fp32 accumulate, fp32 I/O"}[args.math],
                   "launch": tm["launch"], "rank_cores": cores},
        "parity": parity,
        "sustained": {"ms_per_step_median": tm["sustained_ms_per_step_median"], "replays": tm["sustained_replays"],
                      "seconds": tm["sustained_seconds"],
                      "value": world * q / (tm["sustained_ms_per_step_median"] * 1e-3)},
        "e2e": e2e,
        "gpu_launches": int(round(tm["launches_per_step"] * args.steps)),
        "clocks": clocks,
        "roofline": roof,
        "tf32_peak": tf32,
        "cpu_baseline": cpu_block,
        "secondary": secondary,
    }
    emit(line)
    if sampler:
        sampler.stop()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--math", default="tf32", choices=["tf32", "3xtf32"])
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads of the default line")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)
    return run_ours(args, w)


if __name__ == "__main__":
    sys.exit(main())
