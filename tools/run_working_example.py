#!/usr/bin/env python
"""The reference's own training script, UNMODIFIED, end to end on the GPU (working_example.py:88-137):
DECODA loader -> QDNN / QCNN builder (models/example_model.py) -> compile(Adam, categorical_crossentropy) ->
fit(15 epochs, batch 3) -> evaluate, with `complexnn` = this repository's layers and `keras` = the facade.

  python tools/run_working_example.py [--ref DIR] [--model QDNN|QCNN] [--maths tf32,fp32]

For every math mode the script runs from the same seeds (NumPy + torch); the per-epoch curves are printed side by side
(tf32 = tensor-core kernels, fp32 = CUDA-core kernels: the reference's arithmetic), and the trained model of each run is
re-evaluated on the test split by the CPU ORACLE (oracle/qoracle.py forward with the trained weights): its loss /
accuracy must agree with what the GPU's `evaluate` printed.  One JSON line at the end.
The reference checkout is looked up at --ref, else baseline/_ref (a git-ignored copy that travels with gpurun), else
/root/reference."""
import argparse
import io
import json
import os
import runpy
import sys
import time
from contextlib import redirect_stdout

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")


def oracle_eval(model, x, y):
    """Forward of the trained example model on the CPU oracle: the quaternion layers through oracle/qoracle.py, pooling /
    flatten / the real softmax head in NumPy (models/example_model.py:22-47, 69-79)."""
    sys.path.insert(0, REPO)
    from oracle import qoracle as O
    from complexnn.conv import QuaternionConv
    from complexnn.dense import QuaternionDense
    h = np.asarray(x, np.float32)
    for layer in model.layers:
        cls = layer.__class__.__name__
        if isinstance(layer, QuaternionConv):
            ws = layer.get_weights()
            h = O.qconv_forward(h, ws[0], ws[-1] if layer.use_bias else None, layer.filters, layer.strides, layer.padding,
                                layer.data_format, layer.dilation_rate, layer.activation.name)
        elif isinstance(layer, QuaternionDense):
            ws = layer.get_weights()
            h = O.qdense_forward(h, ws[0], ws[1] if layer.use_bias else None, layer.units, layer.activation.name)
        elif cls == "AveragePooling1D":
            pool = layer.pool_size[0] if isinstance(layer.pool_size, (tuple, list)) else layer.pool_size
            n = h.shape[1]
            if layer.padding == "same":        # TF semantics: the divisor counts in-range samples only
                out = -(-n // pool)
                lo = max((out - 1) * pool + pool - n, 0) // 2
            else:
                out, lo = (n - pool) // pool + 1, 0
            h = np.stack([h[:, max(o * pool - lo, 0):min(o * pool - lo + pool, n)].mean(axis=1) for o in range(out)], axis=1)
        elif cls == "Flatten":
            h = h.reshape(h.shape[0], -1)
        elif cls == "Dropout":
            pass
        elif cls == "Dense":
            k, b = layer.get_weights()
            h = h.astype(np.float64) @ k.astype(np.float64) + b
            if layer.activation.name == "softmax":
                e = np.exp(h - h.max(axis=-1, keepdims=True))
                h = e / e.sum(axis=-1, keepdims=True)
            elif layer.activation.name == "relu":
                h = np.maximum(h, 0)
        else:
            raise SystemExit("oracle_eval: unexpected layer " + cls)
    p = np.clip(h, 1e-7, 1 - 1e-7)
    loss = float(-(y * np.log(p)).sum(axis=-1).mean())
    acc = float((h.argmax(-1) == y.argmax(-1)).mean())
    return loss, acc


def run(ref, model, math):
    import torch
    os.environ["QNN_MATH"] = math
    os.environ["QNN_ALGO"] = "general" if math == "fp32" else "auto"
    np.random.seed(0)
    torch.manual_seed(0)
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = ["working_example.py", "--model", model]
    os.chdir(ref)
    buf = io.StringIO()
    t0 = time.time()
    try:
        with redirect_stdout(buf):
            g = runpy.run_path(os.path.join(ref, "working_example.py"), run_name="__main__")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    dt = time.time() - t0
    clf = g["classifier"]
    hist = clf.history.history
    test = clf.evaluate(g["x_test"], g["y_test"])
    o_loss, o_acc = oracle_eval(clf, g["x_test"], g["y_test"])
    tail = [l for l in buf.getvalue().splitlines() if l.startswith("Test Loss")]
    return {"math": math, "seconds": dt, "history": hist, "gpu_test_loss": float(test[0]), "gpu_test_acc": float(test[1]),
            "oracle_test_loss": o_loss, "oracle_test_acc": o_acc, "script_says": tail[-1] if tail else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=None)
    ap.add_argument("--model", default="QDNN", choices=["QDNN", "QCNN"])
    ap.add_argument("--maths", default="tf32,fp32")
    args = ap.parse_args()
    ref = args.ref
    for cand in (os.path.join(REPO, "baseline", "_ref"), "/root/reference"):
        if ref is None and os.path.exists(os.path.join(cand, "working_example.py")):
            ref = cand
    if ref is None:
        raise SystemExit("no reference checkout found (pass --ref, or copy it to baseline/_ref)")
    ref = os.path.abspath(ref)
    sys.path[:0] = [PKG, os.path.join(PKG, "keras_facade")]
    sys.path.append(ref)
    runs = [run(ref, args.model, m) for m in args.maths.split(",")]
    for r in runs:
        h = r["history"]
        print("== %s %s (%.0f s): test loss %.4f acc %.4f | oracle with the trained weights: loss %.4f acc %.4f" % (
            args.model, r["math"], r["seconds"], r["gpu_test_loss"], r["gpu_test_acc"], r["oracle_test_loss"],
            r["oracle_test_acc"]), file=sys.stderr)
        for e in range(len(h["loss"])):
            print("   epoch %2d  loss %.4f acc %.4f  val_loss %.4f val_acc %.4f" % (
                e + 1, h["loss"][e], h.get("acc", [0] * 99)[e], h.get("val_loss", [0] * 99)[e], h.get("val_acc", [0] * 99)[e]),
                file=sys.stderr)
    ok = all(abs(r["gpu_test_loss"] - r["oracle_test_loss"]) <= 2e-3 * max(1.0, abs(r["oracle_test_loss"])) and
             abs(r["gpu_test_acc"] - r["oracle_test_acc"]) <= 0.004 for r in runs)
    print(json.dumps({"script": "working_example.py --model " + args.model, "reference": ref, "runs": runs,
                      "gpu_evaluate_matches_cpu_oracle_on_trained_weights": ok}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
