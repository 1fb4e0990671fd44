// C-ABI entry points of libqnn_b200.so (declared in include/qnn.h): argument validation, geometry resolution
// (TF/Keras padding rules), kernel selection, host-buffer staging and the NCCL gradient exchange.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include "qnn_common.h"

namespace qnn {

std::atomic<unsigned long long> g_launches{0};
static thread_local char t_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

int stream_scratch_alloc(void** ptr, size_t bytes, cudaStream_t st) {
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = 1ull << 30;  // freed blocks stay in the pool: steady-state calls never reach the OS allocator
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    });
    cudaError_t e = cudaMallocAsync(ptr, bytes, st);
    if (e != cudaSuccess) {
        *ptr = nullptr;
        set_error("stream-ordered scratch allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    return QNN_OK;
}

namespace {

bool valid_act(int a) { return a >= QNN_ACT_LINEAR && a <= QNN_ACT_EXPONENTIAL; }

// tf.nn.convolution padding arithmetic (what K.conv1d/2d/3d resolve to; complexnn/conv.py:309-315)
int resolve_axis(int n, int k, int s, int d, int padding, int* pad_lo, int* out) {
    const int eff = (k - 1) * d + 1;
    if (padding == QNN_PAD_VALID) {
        *pad_lo = 0;
        *out = n >= eff ? (n - eff) / s + 1 : 0;
    } else if (padding == QNN_PAD_SAME) {
        *out = (n + s - 1) / s;
        int total = (*out - 1) * s + eff - n;
        if (total < 0) total = 0;
        *pad_lo = total / 2;  // the odd element goes after
    } else if (padding == QNN_PAD_CAUSAL) {
        *out = (n + s - 1) / s;
        *pad_lo = d * (k - 1);
    } else {
        return -1;
    }
    return 0;
}

int build_geom(const qnn_conv_desc* d, Geom* g) {
    if (!d) {
        set_error("descriptor is NULL");
        return QNN_E_INVALID;
    }
    if (d->rank < 1 || d->rank > 3) {
        set_error("rank must be 1, 2 or 3 (got %d)", d->rank);
        return QNN_E_INVALID;
    }
    if (d->batch < 0 || d->in_q < 1 || d->filters < 1) {
        set_error("batch >= 0, in_q >= 1 and filters >= 1 required (got %d, %d, %d)", d->batch, d->in_q, d->filters);
        return QNN_E_INVALID;
    }
    if (d->padding == QNN_PAD_CAUSAL && d->rank != 1) {
        set_error("causal padding is only defined for rank 1");
        return QNN_E_INVALID;
    }
    if (!valid_act(d->activation)) {
        set_error("unknown activation %d", d->activation);
        return QNN_E_INVALID;
    }
    std::memset(g, 0, sizeof(*g));
    g->batch = d->batch;
    g->in_q = d->in_q;
    g->F = d->filters;
    g->channels_first = d->channels_first ? 1 : 0;
    g->act = d->activation;
    g->conj_w = 0;
    const int lead = 3 - d->rank;
    for (int a = 0; a < 3; ++a) {
        g->in_sp[a] = g->out_sp[a] = g->k[a] = g->s[a] = g->d[a] = 1;
        g->pad_lo[a] = 0;
    }
    for (int a = 0; a < d->rank; ++a) {
        const int n = d->in_spatial[a], k = d->kernel[a], s = d->stride[a], dl = d->dilation[a];
        if (n < 0 || k < 1 || s < 1 || dl < 1) {
            set_error("axis %d: need in_spatial >= 0, kernel >= 1, stride >= 1, dilation >= 1 (got %d, %d, %d, %d)", a, n, k,
                      s, dl);
            return QNN_E_INVALID;
        }
        int pad_lo, out;
        if (resolve_axis(n, k, s, dl, d->padding, &pad_lo, &out)) {
            set_error("unknown padding mode %d", d->padding);
            return QNN_E_INVALID;
        }
        g->in_sp[lead + a] = n;
        g->out_sp[lead + a] = out;
        g->k[lead + a] = k;
        g->s[lead + a] = s;
        g->d[lead + a] = dl;
        g->pad_lo[lead + a] = pad_lo;
    }
    return QNN_OK;
}

int build_dense_geom(int64_t rows, int in_q, int q_units, int act, Geom* g) {
    if (rows < 0 || rows > 0x7fffffffLL || in_q < 1 || q_units < 1) {
        set_error("dense: need 0 <= rows < 2^31, in_q >= 1, q_units >= 1 (got %lld, %d, %d)", (long long)rows, in_q,
                  q_units);
        return QNN_E_INVALID;
    }
    if (!valid_act(act)) {
        set_error("unknown activation %d", act);
        return QNN_E_INVALID;
    }
    std::memset(g, 0, sizeof(*g));
    g->batch = 1;
    g->in_q = in_q;
    g->F = q_units;
    g->act = act;
    g->conj_w = 1;
    for (int a = 0; a < 3; ++a) {
        g->in_sp[a] = g->out_sp[a] = g->k[a] = g->s[a] = g->d[a] = 1;
        g->pad_lo[a] = 0;
    }
    g->in_sp[2] = g->out_sp[2] = (int)rows;  // rows are the positions of one channels_last "sequence", one tap
    return QNN_OK;
}

bool empty_out(const Geom& g) { return g.batch == 0 || g.out_sp[0] == 0 || g.out_sp[1] == 0 || g.out_sp[2] == 0; }

int run_forward(const Geom& g, int rank, int math, int algo, const float* x, const float* w, const float* bias, float* y,
                cudaStream_t st) {
    if (empty_out(g)) return QNN_OK;
    if (!x || !w || !y) {
        set_error("x, kernel and y must not be NULL");
        return QNN_E_INVALID;
    }
    if (math != QNN_MATH_TF32 && math != QNN_MATH_FP32 && math != QNN_MATH_3XTF32) {
        set_error("unknown math mode %d", math);
        return QNN_E_INVALID;
    }
    if (algo == QNN_ALGO_GENERAL || math == QNN_MATH_FP32) return general_forward(g, x, w, bias, y, st);
    if (math == QNN_MATH_3XTF32) {
        if (algo == QNN_ALGO_TENSOR) {
            set_error("3xTF32 is not implemented on the tensor-core kernel yet");
            return QNN_E_UNSUPPORTED;
        }
        return general_forward(g, x, w, bias, y, st);
    }
    const bool aligned =
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w)) & 15) == 0;
    if (algo == QNN_ALGO_TENSOR)
        return (g.channels_first || (!tc_plan(g, rank).ok && tc2d_plan(g, rank).ok)) ? tc2d_forward(g, rank, x, w, bias, y, st)
                                                                                      : tc_forward(g, rank, x, w, bias, y, st);
    if (algo != QNN_ALGO_AUTO) {
        set_error("unknown algo %d", algo);
        return QNN_E_INVALID;
    }
    if (aligned && tc_plan(g, rank).ok) return tc_forward(g, rank, x, w, bias, y, st);
    if (aligned && tc2d_plan(g, rank).ok) return tc2d_forward(g, rank, x, w, bias, y, st);
    return general_forward(g, x, w, bias, y, st);
}

// the transposed convolution of the data gradient: dz [batch, Lo, 4F] -> dx [batch, L, 4 in_q]
Geom transposed_geom(const Geom& g) {
    Geom gt = g;
    gt.in_q = g.F;
    gt.F = g.in_q;
    gt.in_sp[2] = g.out_sp[2];
    gt.out_sp[2] = g.in_sp[2];
    gt.pad_lo[2] = (g.k[2] - 1) * g.d[2] - g.pad_lo[2];
    gt.act = QNN_ACT_LINEAR;
    gt.conj_w = g.conj_w ? 0 : 1;
    return gt;
}

// which gradients of this problem the tensor-core kernels take (pointer alignment aside)
void backward_selection(const Geom& g, int rank, int math, int algo, int* dx_tc, int* dw_tc) {
    const bool base = math == QNN_MATH_TF32 && algo != QNN_ALGO_GENERAL && !g.channels_first && rank == 1 &&
                      !empty_out(g) && g.in_sp[2] > 0 && (g.act == QNN_ACT_LINEAR || g.act == QNN_ACT_RELU);
    *dx_tc = base && tc_plan(transposed_geom(g), 1).ok;
    *dw_tc = base && wgrad_plan(g, 1).ok;
}

// Backward.  Tensor-core path (channels_last rank 1 / dense, stride 1): one pass makes dz = dy * act'(y) and the bias
// gradient; the data gradient is the SAME fused Hamilton kernel run on dz with the transposed, tap-flipped stored kernel
// and the transposed sign table; the kernel gradient contracts x with dz over positions on the tensor
// cores, the 16 blocks folding into the 4 stored sub-filters inside tensor memory (qnn_wgrad_tc.cu).  Whatever a
// tensor-core kernel does not take falls to the CUDA-core kernels, piece by piece.
int run_backward(const Geom& g, int rank, int math, int algo, const float* x, const float* w, const float* y,
                 const float* dy, float* dx, float* dw, float* db, cudaStream_t st) {
    if (math != QNN_MATH_TF32 && math != QNN_MATH_FP32 && math != QNN_MATH_3XTF32) {
        set_error("unknown math mode %d", math);
        return QNN_E_INVALID;
    }
    if (algo != QNN_ALGO_AUTO && algo != QNN_ALGO_GENERAL && algo != QNN_ALGO_TENSOR) {
        set_error("unknown algo %d", algo);
        return QNN_E_INVALID;
    }
    const Geom gt = transposed_geom(g);
    auto al16 = [](const void* a) { return (reinterpret_cast<uintptr_t>(a) & 15) == 0; };
    int sel_dx = 0, sel_dw = 0;
    backward_selection(g, rank, math, algo, &sel_dx, &sel_dw);
    const bool al_base = al16(dy) && al16(y) && al16(x);
    const bool tc_dx = sel_dx && al_base && dx && al16(dx);
    const bool tc_dw = sel_dw && al_base && dw && al16(dw);
    if (algo == QNN_ALGO_TENSOR && ((dx && !tc_dx) || (dw && !tc_dw))) {
        set_error("tensor-core backward does not take this problem (channels_last rank 1 / dense, stride 1): dx: %s; "
                  "dkernel: %s", dx ? (tc_dx ? "ok" : tc_plan(gt, 1).why) : "-", dw ? (tc_dw ? "ok" : wgrad_plan(g, 1).why) : "-");
        return QNN_E_UNSUPPORTED;
    }
    if (!tc_dx && !tc_dw) return general_backward(g, x, w, y, dy, dx, dw, db, st);
    const long long rows = (long long)g.batch * g.out_sp[2];
    const int C = 4 * g.F, taps = g.k[2];
    const bool relu = g.act == QNN_ACT_RELU;
    float* dz = nullptr;
    float* wt = nullptr;
    int rc = QNN_OK;
    if (relu && (rc = stream_scratch_alloc(reinterpret_cast<void**>(&dz), (size_t)rows * C * sizeof(float), st))) return rc;
    const float* dzc = relu ? dz : dy;
    Geom gl = g;
    gl.act = QNN_ACT_LINEAR;  // dz already carries the activation derivative
    if (relu || db) rc = dz_bgrad(y, dy, dz, db, rows, C, relu ? 1 : 0, st);
    if (!rc && dx) {
        if (tc_dx) {
            rc = stream_scratch_alloc(reinterpret_cast<void**>(&wt), (size_t)taps * g.in_q * C * sizeof(float), st);
            if (!rc) rc = transpose_w(w, wt, taps, g.in_q, g.F, st);
            if (!rc) rc = tc_forward(gt, 1, dzc, wt, nullptr, dx, st);
        } else {
            rc = general_backward(gl, x, w, dzc, dzc, dx, nullptr, nullptr, st);
        }
    }
    if (!rc && dw) rc = tc_dw ? wgrad_tc(g, 1, x, dzc, dw, st) : general_backward(gl, x, w, dzc, dzc, nullptr, dw, nullptr, st);
    if (dz) cudaFreeAsync(dz, st);
    if (wt) cudaFreeAsync(wt, st);
    return rc;
}

// ---------------------------------------------------------------- device scratch for the *_host entry points
struct Scratch {
    std::mutex mu;
    void* ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t need) {
        if (need <= bytes) return QNN_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&ptr, need);
        if (e != cudaSuccess) {
            set_error("device scratch allocation of %zu bytes failed: %s", need, cudaGetErrorString(e));
            return QNN_E_CUDA;
        }
        bytes = need;
        return QNN_OK;
    }
};
Scratch g_scratch;

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

// Copy engines run next to the SMs and PCIe is full duplex: the units (samples for a convolution, rows for a dense
// layer) are cut into chunks and pipelined over three streams -- H2D of chunk i+1, kernel of chunk i and D2H of chunk
// i-1 overlap -- so the call costs about max(H2D, D2H) instead of their sum.
struct HostPipe {
    cudaStream_t in = nullptr, out = nullptr;
    cudaEvent_t ev_in[16] = {}, ev_k[16] = {}, ev_start = nullptr;
    bool ok = false;
    bool init() {
        if (ok) return true;
        if (cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking) != cudaSuccess) return false;
        for (int i = 0; i < 16; ++i) {
            if (cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        if (cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming) != cudaSuccess) return false;
        ok = true;
        return true;
    }
};
HostPipe g_pipe;

int forward_host(const Geom& g, int rank, int math, int algo, size_t nx, size_t nw, size_t nb, size_t ny,
                 const float* xh, const float* wh, const float* bh, float* yh, cudaStream_t st) {
    if (ny == 0) return QNN_OK;
    if (!xh || !wh || !yh) {
        set_error("host x, kernel and y must not be NULL");
        return QNN_E_INVALID;
    }
    std::lock_guard<std::mutex> lock(g_scratch.mu);
    const size_t ox = 0, ow = ox + align256(nx * 4), ob = ow + align256(nw * 4), oy = ob + align256(nb * 4);
    int rc = g_scratch.reserve(oy + align256(ny * 4) + 256);
    if (rc) return rc;
    char* base = static_cast<char*>(g_scratch.ptr);
    float *xd = (float*)(base + ox), *wd = (float*)(base + ow), *bd = bh ? (float*)(base + ob) : nullptr,
          *yd = (float*)(base + oy);
    // units that can be cut independently: samples (conv) or rows (dense: batch == 1, rows live in in_sp[2])
    const bool dense = g.conj_w && g.batch == 1 && g.k[2] == 1;
    const long long units = dense ? g.in_sp[2] : g.batch;
    const size_t x_unit = nx / (size_t)units, y_unit = ny / (size_t)units;
    int chunks = 1;
    if ((nx + ny) * 4 >= (size_t)8 << 20 && units >= 8 && g_pipe.init()) {
        chunks = units >= 64 ? 8 : 4;  // measured on cfg 2: 1 / 2 / 4 / 8 / 16 chunks -> 2.04 / 1.70 / 1.59 / 1.54 / 1.65 ms
                                       // (67 MB of y over PCIe is ~1.45 ms on its own; uneven cuts gained nothing)
        if (const char* env = getenv("QNN_HOST_CHUNKS")) {  // tuning knob: 1..16 pipeline chunks
            const int v = atoi(env);
            if (v >= 1 && v <= 16 && v <= units) chunks = v;
        }
    }
    cudaError_t e;
    if ((e = cudaMemcpyAsync(wd, wh, nw * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) goto fail;
    if (bh && (e = cudaMemcpyAsync(bd, bh, nb * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) goto fail;
    if (chunks == 1) {
        if ((e = cudaMemcpyAsync(xd, xh, nx * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) goto fail;
        rc = run_forward(g, rank, math, algo, xd, wd, bd, yd, st);
        if (rc) return rc;
        if ((e = cudaMemcpyAsync(yh, yd, ny * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) goto fail;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) goto fail;
        return QNN_OK;
    }
    // the copy streams must not run ahead of work already queued on the caller's stream (scratch reuse)
    if ((e = cudaEventRecord(g_pipe.ev_start, st)) != cudaSuccess) goto fail;
    if ((e = cudaStreamWaitEvent(g_pipe.in, g_pipe.ev_start, 0)) != cudaSuccess) goto fail;
    if ((e = cudaStreamWaitEvent(g_pipe.out, g_pipe.ev_start, 0)) != cudaSuccess) goto fail;
    for (int c = 0; c < chunks; ++c) {
        const long long u0 = units * c / chunks, u1 = units * (c + 1) / chunks;
        if (u1 == u0) continue;
        Geom gc = g;
        if (dense)
            gc.in_sp[2] = gc.out_sp[2] = (int)(u1 - u0);
        else
            gc.batch = (int)(u1 - u0);
        if ((e = cudaMemcpyAsync(xd + u0 * x_unit, xh + u0 * x_unit, (size_t)(u1 - u0) * x_unit * 4,
                                 cudaMemcpyHostToDevice, g_pipe.in)) != cudaSuccess) goto fail;
        if ((e = cudaEventRecord(g_pipe.ev_in[c], g_pipe.in)) != cudaSuccess) goto fail;
        if ((e = cudaStreamWaitEvent(st, g_pipe.ev_in[c], 0)) != cudaSuccess) goto fail;
        rc = run_forward(gc, rank, math, algo, xd + u0 * x_unit, wd, bd, yd + u0 * y_unit, st);
        if (rc) return rc;
        if ((e = cudaEventRecord(g_pipe.ev_k[c], st)) != cudaSuccess) goto fail;
        if ((e = cudaStreamWaitEvent(g_pipe.out, g_pipe.ev_k[c], 0)) != cudaSuccess) goto fail;
        if ((e = cudaMemcpyAsync(yh + u0 * y_unit, yd + u0 * y_unit, (size_t)(u1 - u0) * y_unit * 4,
                                 cudaMemcpyDeviceToHost, g_pipe.out)) != cudaSuccess) goto fail;
    }
    if ((e = cudaStreamSynchronize(g_pipe.out)) != cudaSuccess) goto fail;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) goto fail;
    return QNN_OK;
fail:
    set_error("host staging failed: %s", cudaGetErrorString(e));
    return QNN_E_CUDA;
}

// ---------------------------------------------------------------- NCCL through dlopen (no link-time dependency)
typedef struct ncclComm* ncclComm_t;
struct NcclUid {  // ncclUniqueId: 128 opaque bytes, passed by value
    char b[128];
};
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUid, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
ncclComm_t g_comm = nullptr;
std::mutex g_comm_mu;

int load_nccl() {
    if (g_nccl.lib) return QNN_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)
        if ((h = dlopen(n, RTLD_NOW | RTLD_NOLOAD))) break;  // prefer the copy already in the process (torch's)
    if (!h)
        for (const char* n : names)
            if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) {
        set_error("NCCL not found: %s", dlerror());
        return QNN_E_COMM;
    }
    g_nccl.lib = h;
    *(void**)&g_nccl.GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void**)&g_nccl.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void**)&g_nccl.AllReduce = dlsym(h, "ncclAllReduce");
    *(void**)&g_nccl.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void**)&g_nccl.GetErrorString = dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        set_error("NCCL library lacks a required symbol");
        g_nccl.lib = nullptr;
        return QNN_E_COMM;
    }
    return QNN_OK;
}

int nccl_fail(const char* what, int code) {
    set_error("%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(code) : "nccl error");
    return QNN_E_COMM;
}

}  // namespace
}  // namespace qnn

using namespace qnn;

extern "C" {

int qnn_abi_version(void) { return QNN_ABI_VERSION; }
const char* qnn_last_error(void) { return t_err; }
uint64_t qnn_launch_count(void) { return g_launches.load(); }

int qnn_conv_out_spatial(const qnn_conv_desc* d, int32_t out_spatial[3]) {
    Geom g;
    int rc = build_geom(d, &g);
    if (rc) return rc;
    if (!out_spatial) {
        set_error("out_spatial is NULL");
        return QNN_E_INVALID;
    }
    for (int a = 0; a < 3; ++a) out_spatial[a] = a < d->rank ? g.out_sp[3 - d->rank + a] : 1;
    return QNN_OK;
}

int qnn_conv_uses_tensor_cores(const qnn_conv_desc* d) {
    Geom g;
    if (build_geom(d, &g)) return 0;
    if (d->algo == QNN_ALGO_GENERAL || d->math != QNN_MATH_TF32) return 0;
    return tc_plan(g, d->rank).ok || tc2d_plan(g, d->rank).ok;
}

int qnn_dense_uses_tensor_cores(int64_t rows, int32_t in_q, int32_t q_units) {
    Geom g;
    if (build_dense_geom(rows, in_q, q_units, QNN_ACT_LINEAR, &g)) return 0;
    return tc_plan(g, 1).ok;
}

int qnn_conv_backward_uses_tensor_cores(const qnn_conv_desc* d, int32_t* dx_tc, int32_t* dkernel_tc) {
    Geom g;
    int rc = build_geom(d, &g);
    if (rc) return rc;
    if (!dx_tc || !dkernel_tc) {
        set_error("output pointers must not be NULL");
        return QNN_E_INVALID;
    }
    int a = 0, b = 0;
    backward_selection(g, d->rank, d->math, d->algo, &a, &b);
    *dx_tc = a;
    *dkernel_tc = b;
    return QNN_OK;
}

int qnn_dense_backward_uses_tensor_cores(int64_t rows, int32_t in_q, int32_t q_units, int32_t* dx_tc, int32_t* dkernel_tc) {
    Geom g;
    int rc = build_dense_geom(rows, in_q, q_units, QNN_ACT_LINEAR, &g);
    if (rc) return rc;
    if (!dx_tc || !dkernel_tc) {
        set_error("output pointers must not be NULL");
        return QNN_E_INVALID;
    }
    int a = 0, b = 0;
    backward_selection(g, 1, QNN_MATH_TF32, QNN_ALGO_AUTO, &a, &b);
    *dx_tc = a;
    *dkernel_tc = b;
    return QNN_OK;
}

int qnn_conv_forward(const qnn_conv_desc* d, const float* x, const float* kernel, const float* bias, float* y,
                     void* stream) {
    Geom g;
    int rc = build_geom(d, &g);
    if (rc) return rc;
    return run_forward(g, d->rank, d->math, d->algo, x, kernel, bias, y, static_cast<cudaStream_t>(stream));
}

int qnn_dense_forward(int64_t rows, int32_t in_q, int32_t q_units, const float* x, const float* kernel,
                      const float* bias, int32_t activation, int32_t math, int32_t algo, float* y, void* stream) {
    Geom g;
    int rc = build_dense_geom(rows, in_q, q_units, activation, &g);
    if (rc) return rc;
    return run_forward(g, 1, math, algo, x, kernel, bias, y, static_cast<cudaStream_t>(stream));
}

int qnn_conv_backward(const qnn_conv_desc* d, const float* x, const float* kernel, const float* y, const float* dy,
                      float* dx, float* dkernel, float* dbias, void* stream) {
    Geom g;
    int rc = build_geom(d, &g);
    if (rc) return rc;
    if (g.act != QNN_ACT_LINEAR && g.act != QNN_ACT_RELU) {
        set_error("backward supports linear and relu activations only");
        return QNN_E_UNSUPPORTED;
    }
    if (!x || !kernel || !y || !dy) {
        if (empty_out(g) && !dkernel && !dbias && !dx) return QNN_OK;
        set_error("x, kernel, y and dy must not be NULL");
        return QNN_E_INVALID;
    }
    return run_backward(g, d->rank, d->math, d->algo, x, kernel, y, dy, dx, dkernel, dbias,
                        static_cast<cudaStream_t>(stream));
}

int qnn_dense_backward(int64_t rows, int32_t in_q, int32_t q_units, const float* x, const float* kernel,
                       const float* y, const float* dy, int32_t activation, int32_t math, int32_t algo, float* dx,
                       float* dkernel, float* dbias, void* stream) {
    Geom g;
    int rc = build_dense_geom(rows, in_q, q_units, activation, &g);
    if (rc) return rc;
    if (g.act != QNN_ACT_LINEAR && g.act != QNN_ACT_RELU) {
        set_error("backward supports linear and relu activations only");
        return QNN_E_UNSUPPORTED;
    }
    if (!x || !kernel || !y || !dy) {
        set_error("x, kernel, y and dy must not be NULL");
        return QNN_E_INVALID;
    }
    return run_backward(g, 1, math, algo, x, kernel, y, dy, dx, dkernel, dbias, static_cast<cudaStream_t>(stream));
}

int qnn_conv_forward_host(const qnn_conv_desc* d, const float* x_host, const float* kernel_host,
                          const float* bias_host, float* y_host, void* stream) {
    Geom g;
    int rc = build_geom(d, &g);
    if (rc) return rc;
    const size_t S = (size_t)g.in_sp[0] * g.in_sp[1] * g.in_sp[2], P = (size_t)g.out_sp[0] * g.out_sp[1] * g.out_sp[2];
    const size_t taps = (size_t)g.k[0] * g.k[1] * g.k[2];
    return forward_host(g, d->rank, d->math, d->algo, (size_t)g.batch * S * 4 * g.in_q, taps * g.in_q * 4 * g.F,
                        (size_t)4 * g.F, (size_t)g.batch * P * 4 * g.F, x_host, kernel_host, bias_host, y_host,
                        static_cast<cudaStream_t>(stream));
}

int qnn_dense_forward_host(int64_t rows, int32_t in_q, int32_t q_units, const float* x_host, const float* kernel_host,
                           const float* bias_host, int32_t activation, int32_t math, int32_t algo, float* y_host,
                           void* stream) {
    Geom g;
    int rc = build_dense_geom(rows, in_q, q_units, activation, &g);
    if (rc) return rc;
    return forward_host(g, 1, math, algo, (size_t)rows * 4 * in_q, (size_t)in_q * 4 * q_units, (size_t)4 * q_units,
                        (size_t)rows * 4 * q_units, x_host, kernel_host, bias_host, y_host,
                        static_cast<cudaStream_t>(stream));
}

int qnn_debug_trace(void* device_buffer, size_t bytes) {
    tc_set_trace(device_buffer, bytes);
    tc2d_set_trace(device_buffer, bytes);
    wgrad_set_trace(device_buffer, bytes);
    return QNN_OK;
}

int qnn_comm_unique_id(void* out_128_bytes) {
    std::lock_guard<std::mutex> lock(g_comm_mu);
    if (!out_128_bytes) {
        set_error("output buffer is NULL");
        return QNN_E_INVALID;
    }
    int rc = load_nccl();
    if (rc) return rc;
    int e = g_nccl.GetUniqueId(out_128_bytes);
    return e ? nccl_fail("ncclGetUniqueId", e) : QNN_OK;
}

int qnn_comm_init(int32_t rank, int32_t world_size, const void* unique_id_128_bytes) {
    std::lock_guard<std::mutex> lock(g_comm_mu);
    if (world_size < 1 || rank < 0 || rank >= world_size || !unique_id_128_bytes) {
        set_error("need 0 <= rank < world_size and a unique id");
        return QNN_E_INVALID;
    }
    if (g_comm) {
        set_error("communicator already initialised");
        return QNN_E_STATE;
    }
    int rc = load_nccl();
    if (rc) return rc;
    NcclUid uid;
    std::memcpy(uid.b, unique_id_128_bytes, 128);
    int e = g_nccl.CommInitRank(&g_comm, world_size, uid, rank);
    if (e) {
        g_comm = nullptr;
        return nccl_fail("ncclCommInitRank", e);
    }
    return QNN_OK;
}

int qnn_allreduce_f32(float* buf, size_t count, void* stream) {
    std::lock_guard<std::mutex> lock(g_comm_mu);
    if (!g_comm) {
        set_error("qnn_comm_init has not been called");
        return QNN_E_STATE;
    }
    if (count == 0) return QNN_OK;
    if (!buf) {
        set_error("buffer is NULL");
        return QNN_E_INVALID;
    }
    int e = g_nccl.AllReduce(buf, buf, count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, g_comm, static_cast<cudaStream_t>(stream));
    return e ? nccl_fail("ncclAllReduce", e) : QNN_OK;
}

int qnn_comm_destroy(void) {
    std::lock_guard<std::mutex> lock(g_comm_mu);
    if (!g_comm) return QNN_OK;
    int e = g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
    return e ? nccl_fail("ncclCommDestroy", e) : QNN_OK;
}

}  // extern "C"
