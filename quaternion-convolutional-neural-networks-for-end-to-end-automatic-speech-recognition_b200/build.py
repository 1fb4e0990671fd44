"""Builds libqnn_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.  No torch, no libcuda link."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("QNN_BUILD_LIB") or os.path.join(HERE, "lib", "libqnn_b200.so")  # QNN_BUILD_LIB: build a variant elsewhere
SOURCES = ["qnn_api.cu", "qnn_general.cu", "qnn_smallk.cu", "qnn_hamilton_tc.cu", "qnn_hamilton_tc2d.cu", "qnn_wgrad_tc.cu"]
HEADERS = ["qnn_common.h", "qnn_ptx.cuh", "qnn_tmap.h", os.path.join("..", "..", "include", "qnn.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v"]
# --use_fast_math (flush-to-zero, approximate division / transcendentals) only where it is harmless: the tensor-core
# kernels' epilogues.  qnn_general.cu is the FP32-faithful path the tests use as a reference: IEEE arithmetic there.
FAST_MATH = {"qnn_hamilton_tc.cu", "qnn_hamilton_tc2d.cu", "qnn_wgrad_tc.cu"}


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = list(NVCC_FLAGS)
    if os.environ.get("QNN_NVCC_DEFINES"):  # experiment builds: e.g. QNN_NVCC_DEFINES="-DQNN_SLEEP_NS=32 -DQNN_WCHUNK_BARS=1"
        flags += os.environ["QNN_NVCC_DEFINES"].split()
    if os.environ.get("QNN_SPIN_LIMIT"):  # debug builds: bounded mbarrier waits that trap instead of hanging the GPU
        flags.append("-DQNN_SPIN_LIMIT=" + os.environ["QNN_SPIN_LIMIT"])
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(os.path.dirname(LIB), s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc] + flags + (["--use_fast_math"] if s in FAST_MATH else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(" ".join(cmd))
            print(out)
        if p.returncode:
            raise RuntimeError("nvcc failed for " + cmd[-3])
    link = [nvcc, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl", "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode:
        print(" ".join(link))
        print(r.stdout)
    if r.returncode:
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
