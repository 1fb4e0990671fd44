// C-ABI entry points of libqnn_b200.so (declared in include/qnn.h): argument validation, geometry resolution
// (TF/Keras padding rules), kernel selection, host-buffer staging and the NCCL gradient exchange.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>
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

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return dev;
}

int num_sms() {
    static std::atomic<int> n[kMaxDevices];
    const int dev = current_device();
    int v = n[dev].load(std::memory_order_relaxed);
    if (!v) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// cudaFuncSetAttribute is per device: remember (device, kernel) pairs, not kernels
int ensure_dynamic_smem(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::vector<std::pair<int, const void*>> done;
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(mu);
    for (const auto& d : done)
        if (d.first == dev && d.second == kernel) return QNN_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize = %d) failed: %s", bytes, cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    done.emplace_back(dev, kernel);
    return QNN_OK;
}

int stream_scratch_alloc(void** ptr, size_t bytes, cudaStream_t st) {
    // every device's default pool is told once to keep freed blocks: steady-state calls never reach the OS allocator
    static std::atomic<bool> pool_ready[kMaxDevices];
    const int dev = current_device();
    if (!pool_ready[dev].load(std::memory_order_acquire)) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = 1ull << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool_ready[dev].store(true, std::memory_order_release);
    }
    cudaError_t e = cudaMallocAsync(ptr, bytes ? bytes : 16, st);
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
    g->dense = 1;
    for (int a = 0; a < 3; ++a) {
        g->in_sp[a] = g->out_sp[a] = g->k[a] = g->s[a] = g->d[a] = 1;
        g->pad_lo[a] = 0;
    }
    g->in_sp[2] = g->out_sp[2] = (int)rows;  // rows are the positions of one channels_last "sequence", one tap
    return QNN_OK;
}

bool empty_out(const Geom& g) { return g.batch == 0 || g.out_sp[0] == 0 || g.out_sp[1] == 0 || g.out_sp[2] == 0; }

bool valid_math(int m) { return m == QNN_MATH_TF32 || m == QNN_MATH_FP32 || m == QNN_MATH_3XTF32; }
bool valid_algo(int a) { return a == QNN_ALGO_AUTO || a == QNN_ALGO_GENERAL || a == QNN_ALGO_TENSOR; }
bool al16(const void* a) { return (reinterpret_cast<uintptr_t>(a) & 15) == 0; }

// Which kernel a forward problem gets: 0 = general (CUDA cores), 1 = channels_last rows (qnn_hamilton_tc.cu),
// 2 = channels_first / rank 2 (qnn_hamilton_tc2d.cu), 3 = small-K (CUDA cores with shuffle tap reuse, qnn_smallk.cu:
// fp32 FMA, so it serves every math mode).  Pointer alignment aside.
enum { kKernGeneral = 0, kKernTc = 1, kKernTc2d = 2, kKernSmallK = 3 };
int forward_kernel(const Geom& g, int rank, int math, int algo) {
    if (algo == QNN_ALGO_GENERAL) return kKernGeneral;
    if (algo == QNN_ALGO_AUTO && smallk_plan(g, rank).ok) return kKernSmallK;
    if (math == QNN_MATH_FP32) return kKernGeneral;
    const int x3 = math == QNN_MATH_3XTF32;
    // Resident sub-filters (one launch keeps the whole image in shared memory) win when the image fits with 64-filter
    // tiles or the layer is one tile; when it only fits with 32-filter tiles the kernel makes twice the passes over x with
    // N = 32 MMAs that cost as much as N = 64 ones (measured: cfg 3 conv layers 60 us vs ~35 us) -- the streamed-sub-filter
    // kernel then takes the problem if it qualifies (in_q % 8 == 0, F % 32 == 0).
    // A resident image that leaves room for fewer than four x stages starves the converters: with two stages every
    // converter group has ONE x buffer and waits out a whole load per stage (the TIMIT first layer of the cfg 3 stack,
    // in_q = 41, F = 64, k = 3, 147 KB of image: 23 k cycles per tile against 11 k of MMAs).  The streamed-sub-filter kernel
    // (six x stages, ragged rows read in place) takes such layers when it needs no pre-pass -- measured on one box
    // (profiles/r02_ab_starved_layers.txt): in_q = 41: stack -9.5 %; in_q = 44 / 48 / 96 (F = 32): -22 / -10 / -15 %.
    // QNN_STARVED_STREAM=0 keeps them on the resident kernel (A/B timing).
    static const bool starved_stream = [] { const char* e = getenv("QNN_STARVED_STREAM"); return !(e && atoi(e) == 0); }();
    const TcPlan tcp = tc_plan(g, rank, x3);
    const Tc2dPlan t2p = tc2d_plan(g, rank, x3);
    const bool ragged_starved = starved_stream && tcp.ok && tcp.x_stages < 4 && t2p.ok && !t2p.pad_q && !t2p.pad_rows;
    // (QNN_FORCE_STREAM=1, experiments: the streamed kernel whenever it takes the problem)
    static const bool force_stream = [] { const char* e = getenv("QNN_FORCE_STREAM"); return e && atoi(e) != 0; }();
    // More than one filter-tile pass (F > 64): the resident kernel re-reads x and reloads an image per pass; the streamed
    // kernel is 1.5-3 % faster at two passes and 10 % at four (profiles/r02_ab_starved_layers.txt, second block)
    const bool multi_pass = tcp.ok && tcp.n_ftiles > 1 && t2p.ok && !t2p.pad_q && !t2p.pad_rows && starved_stream;
    const bool resident_good = tcp.ok && (tcp.n_ftiles == 1 || tcp.f_tile >= 64) && !ragged_starved && !multi_pass &&
                               !(force_stream && t2p.ok && !t2p.pad_q);
    if (resident_good) return kKernTc;
    if (t2p.ok) return kKernTc2d;
    if (tcp.ok) return kKernTc;
    return kKernGeneral;
}

// `packed`: the caller's cached kernel image (qnn_conv_pack / qnn_dense_pack) or NULL (pack per call into scratch).
int run_forward(const Geom& g, int rank, int math, int algo, const float* x, const float* w, const void* packed,
                const float* bias, float* y, cudaStream_t st) {
    if (!valid_math(math)) {
        set_error("unknown math mode %d", math);
        return QNN_E_INVALID;
    }
    if (!valid_algo(algo)) {
        set_error("unknown algo %d", algo);
        return QNN_E_INVALID;
    }
    if (empty_out(g)) return QNN_OK;
    if (!x || (!w && !packed) || !y) {
        set_error("x, kernel and y must not be NULL");
        return QNN_E_INVALID;
    }
    const int x3 = math == QNN_MATH_3XTF32;
    int kern = forward_kernel(g, rank, math, algo);
    const bool aligned = al16(x) && al16(y) && (packed ? al16(packed) : al16(w));
    if (algo == QNN_ALGO_TENSOR) {
        if (math == QNN_MATH_FP32) {
            set_error("the tensor-core kernels compute in TF32 or 3xTF32, not FP32");
            return QNN_E_UNSUPPORTED;
        }
        if (kern == kKernGeneral) kern = g.channels_first ? kKernTc2d : kKernTc;  // let the kernel's own plan report why not
    } else if (!aligned) {
        kern = kKernGeneral;
    }
    if (kern == kKernSmallK) {
        if (!w) {
            set_error("this problem runs on the small-K kernel, which needs the stored kernel (not only its packed image)");
            return QNN_E_INVALID;
        }
        return smallk_forward(g, rank, x, w, bias, y, st);
    }
    if (kern == kKernGeneral) {
        if (!w) {
            set_error("this problem runs on the general kernel, which needs the stored kernel (not only its packed image)");
            return QNN_E_INVALID;
        }
        return general_forward(g, x, w, bias, y, st);
    }
    if (kern == kKernTc)
        return packed ? tc_forward_packed(g, rank, x3, x, packed, bias, y, st) : tc_forward(g, rank, x3, 0, x, w, bias, y, st);
    return packed ? tc2d_forward_packed(g, rank, x3, x, packed, bias, y, st) : tc2d_forward(g, rank, x3, 0, x, w, bias, y, st);
}

// the transposed convolution of the data gradient: dz [batch, out..., 4F] -> dx [batch, in..., 4 in_q]
Geom transposed_geom(const Geom& g) {
    Geom gt = g;
    gt.in_q = g.F;
    gt.F = g.in_q;
    for (int a = 0; a < 3; ++a) {
        gt.in_sp[a] = g.out_sp[a];
        gt.out_sp[a] = g.in_sp[a];
        gt.pad_lo[a] = (g.k[a] - 1) * g.d[a] - g.pad_lo[a];
    }
    gt.act = QNN_ACT_LINEAR;
    gt.conj_w = g.conj_w ? 0 : 1;
    return gt;
}

// which gradients of this problem the tensor-core kernels take (pointer alignment aside); *dx_kern = kKern* of the
// transposed problem
void backward_selection(const Geom& g, int rank, int math, int algo, int* dx_kern, int* dw_tc) {
    *dx_kern = kKernGeneral;
    *dw_tc = 0;
    if (math == QNN_MATH_FP32 || algo == QNN_ALGO_GENERAL || empty_out(g)) return;
    if (g.act != QNN_ACT_LINEAR && g.act != QNN_ACT_RELU) return;
    for (int a = 0; a < 3; ++a)
        if (g.s[a] != 1 || g.in_sp[a] < 1) return;   // a strided forward is a dilated-input transposed problem: general kernels
    const int x3 = math == QNN_MATH_3XTF32;
    const Geom gt = transposed_geom(g);
    // the transposed problem must reproduce the input extent exactly (true for stride 1)
    *dx_kern = forward_kernel(gt, rank, math, QNN_ALGO_TENSOR);  // (TENSOR: never the small-K kernel, which has no transposed form)
    *dw_tc = wgrad_plan(g, rank, x3).ok;
}

// Backward.  Tensor-core path (stride 1): one pass makes dz = dy * act'(y) and the bias gradient; the data gradient is
// the SAME fused Hamilton kernel run on dz with the transposed, tap-flipped kernel image (packed straight from the stored
// kernel) and the transposed sign table; the kernel gradient contracts x with dz over positions on the
// tensor cores, the 16 blocks folding into the 4 stored sub-filters inside tensor memory (qnn_wgrad_tc.cu).  Whatever
// a tensor-core kernel does not take falls to the CUDA-core kernels, piece by piece.
int run_backward(const Geom& g, int rank, int math, int algo, const float* x, const float* w, const void* packed_dgrad,
                 const float* y, const float* dy, float* dx, float* dw, float* db, cudaStream_t st) {
    if (!valid_math(math)) {
        set_error("unknown math mode %d", math);
        return QNN_E_INVALID;
    }
    if (!valid_algo(algo)) {
        set_error("unknown algo %d", algo);
        return QNN_E_INVALID;
    }
    const Geom gt = transposed_geom(g);
    const int x3 = math == QNN_MATH_3XTF32;
    int sel_dx = 0, sel_dw = 0;
    backward_selection(g, rank, math, algo, &sel_dx, &sel_dw);
    const bool al_base = al16(dy) && al16(y) && al16(x);
    const int tc_dx = (sel_dx && al_base && dx && al16(dx) && (packed_dgrad ? al16(packed_dgrad) : al16(w))) ? sel_dx : 0;
    const bool tc_dw = sel_dw && al_base && dw && al16(dw);
    if (algo == QNN_ALGO_TENSOR && ((dx && !tc_dx) || (dw && !tc_dw))) {
        set_error("tensor-core backward does not take this problem: dx: %s; dkernel: %s",
                  dx ? (tc_dx ? "ok" : (g.channels_first ? tc2d_plan(gt, rank, x3).why : tc_plan(gt, rank, x3).why)) : "-",
                  dw ? (tc_dw ? "ok" : wgrad_plan(g, rank, x3).why) : "-");
        return QNN_E_UNSUPPORTED;
    }
    if (!tc_dx && !tc_dw) return general_backward(g, x, w, y, dy, dx, dw, db, st);
    const long long P = (long long)g.out_sp[0] * g.out_sp[1] * g.out_sp[2];
    const long long rows = (long long)g.batch * P;
    const int C = 4 * g.F;
    const bool relu = g.act == QNN_ACT_RELU;
    float* dz = nullptr;
    int rc = QNN_OK;
    // relu layer, channels_last rank 1 / dense, kernel gradient on the tensor cores: the kernel-gradient kernel reads dy
    // and y itself, forms dz = relu'(y) * dy in its packers, writes it out for the data gradient and accumulates the bias
    // gradient -- no separate pass over y and dy (it was 28 % of the training step).
    const bool fused = relu && tc_dw && !g.channels_first && g.k[0] == 1 && g.k[1] == 1 && al16(y);
    const bool need_dz = relu && (!fused || dx);
    if (need_dz && (rc = stream_scratch_alloc(reinterpret_cast<void**>(&dz), (size_t)rows * C * sizeof(float), st))) return rc;
    const float* dzc = relu ? dz : dy;
    Geom gl = g;
    gl.act = QNN_ACT_LINEAR;  // dz already carries the activation derivative
    if (fused) {
        rc = wgrad_tc(g, rank, x3, x, dy, dw, st, y, dz, db);
        dw = nullptr;  // done
    } else if (relu || db) {
        rc = g.channels_first ? dz_bgrad_cf(y, dy, dz, db, g.batch, C, P, relu ? 1 : 0, st)
                              : dz_bgrad(y, dy, dz, db, rows, C, relu ? 1 : 0, st);
    }
    if (!rc && dx) {
        if (tc_dx == kKernTc)
            rc = packed_dgrad ? tc_forward_packed(gt, rank, x3, dzc, packed_dgrad, nullptr, dx, st)
                              : tc_forward(gt, rank, x3, 1, dzc, w, nullptr, dx, st);
        else if (tc_dx == kKernTc2d)
            rc = packed_dgrad ? tc2d_forward_packed(gt, rank, x3, dzc, packed_dgrad, nullptr, dx, st)
                              : tc2d_forward(gt, rank, x3, 1, dzc, w, nullptr, dx, st);
        else
            rc = general_backward(gl, x, w, dzc, dzc, dx, nullptr, nullptr, st);
    }
    if (!rc && dw) {
        if (!tc_dw) {
            rc = general_backward(gl, x, w, dzc, dzc, nullptr, dw, nullptr, st);
        } else if (!g.channels_first) {
            rc = wgrad_tc(g, rank, x3, x, dzc, dw, st);
        } else {
            // the kernel-gradient kernel walks channels_last rows: transposed scratch copies of x and dz (two extra passes
            // each; still two orders of magnitude faster than the CUDA-core kernel at TIMIT sizes)
            const long long S = (long long)g.in_sp[0] * g.in_sp[1] * g.in_sp[2];
            float *xt = nullptr, *dzt = nullptr;
            rc = stream_scratch_alloc(reinterpret_cast<void**>(&xt), (size_t)g.batch * S * 4 * g.in_q * sizeof(float), st);
            if (!rc) rc = stream_scratch_alloc(reinterpret_cast<void**>(&dzt), (size_t)rows * C * sizeof(float), st);
            if (!rc) rc = cf_to_cl(x, xt, g.batch, 4 * g.in_q, S, st);
            if (!rc) rc = cf_to_cl(dzc, dzt, g.batch, C, P, st);
            if (!rc) rc = wgrad_tc(g, rank, x3, xt, dzt, dw, st);
            if (xt) cudaFreeAsync(xt, st);
            if (dzt) cudaFreeAsync(dzt, st);
        }
    }
    if (dz) cudaFreeAsync(dz, st);
    return rc;
}

// ---------------------------------------------------------------- the *_host entry points
size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

// Copy engines run next to the SMs and PCIe is full duplex: the units (samples for a convolution, rows for a dense
// layer) are cut into chunks and pipelined over three streams -- H2D of chunk i+1, kernel of chunk i and D2H of chunk
// i-1 overlap -- so the call costs about max(H2D, D2H) instead of their sum.  Everything a call needs is per device
// and per call: a HostPipe (two copy streams + events) is borrowed from the device's free list, staging memory is
// stream-ordered scratch (no cudaMalloc / cudaFree, hence no implicit device synchronisation), so concurrent callers
// on different streams or devices do not serialise on each other.
struct HostPipe {
    cudaStream_t in = nullptr, out = nullptr;
    cudaEvent_t ev_in[16] = {}, ev_k[16] = {}, ev_start = nullptr, ev_out = nullptr;
    bool init() {
        if (cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking) != cudaSuccess) return false;
        for (int i = 0; i < 16; ++i) {
            if (cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        if (cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming) != cudaSuccess) return false;
        return true;
    }
};
struct PipePool {
    std::mutex mu;
    std::vector<HostPipe*> idle;
};
PipePool g_pipes[kMaxDevices];

HostPipe* borrow_pipe(int dev) {
    {
        std::lock_guard<std::mutex> lock(g_pipes[dev].mu);
        if (!g_pipes[dev].idle.empty()) {
            HostPipe* p = g_pipes[dev].idle.back();
            g_pipes[dev].idle.pop_back();
            return p;
        }
    }
    HostPipe* p = new HostPipe();
    if (!p->init()) {  // (a half-built pipe is leaked: creation only fails when the context is already unusable)
        return nullptr;
    }
    return p;
}
void return_pipe(int dev, HostPipe* p) {
    std::lock_guard<std::mutex> lock(g_pipes[dev].mu);
    g_pipes[dev].idle.push_back(p);
}

// Weights uploaded by the *_host entry points stay resident on the device, keyed by (host pointer, size, content hash):
// a caller that passes the same kernel / bias again (every inference step) pays a ~10 us hash instead of an upload
// and -- on the tensor-core path -- instead of a re-pack, since the packed image is kept next to it.
struct ResidentWeights {
    const void* host = nullptr;
    size_t bytes = 0;
    uint64_t hash = 0, tick = 0;
    float* dev_raw = nullptr;
    void* dev_packed = nullptr;  // image for (pack_sig): 0 = none
    uint64_t pack_sig = 0;
    cudaEvent_t last_use = nullptr;
};
struct WeightCache {
    std::mutex mu;
    std::vector<ResidentWeights> entries;
    uint64_t tick = 0;
};
WeightCache g_weights[kMaxDevices];
constexpr size_t kMaxResident = 32;

uint64_t hash_bytes(const void* p, size_t n) {  // FNV-1a over 8-byte words: ~1 GB/s per ... plenty for <= MB-sized kernels
    const uint64_t* w = static_cast<const uint64_t*>(p);
    uint64_t h = 1469598103934665603ull;
    const size_t n8 = n / 8;
    for (size_t i = 0; i < n8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    const unsigned char* t = static_cast<const unsigned char*>(p) + n8 * 8;
    for (size_t i = 0; i < n % 8; ++i) h = (h ^ t[i]) * 1099511628211ull;
    return h;
}

// Returns the device copy of `host` (uploading it on `st` when new or changed).  *entry_idx identifies the entry for
// attach_packed / mark_used.  Entries are recycled least-recently-used; a recycled buffer is freed stream-ordered after
// the last kernel that used it (its last_use event).
int resident_weights(int dev, const float* host, size_t n_floats, cudaStream_t st, float** out, size_t* entry_idx) {
    WeightCache& wc = g_weights[dev];
    const size_t bytes = n_floats * sizeof(float);
    const uint64_t h = hash_bytes(host, bytes);
    std::lock_guard<std::mutex> lock(wc.mu);
    for (size_t i = 0; i < wc.entries.size(); ++i) {
        ResidentWeights& e = wc.entries[i];
        if (e.host == host && e.bytes == bytes && e.hash == h) {
            e.tick = ++wc.tick;
            cudaStreamWaitEvent(st, e.last_use, 0);  // another stream may still be uploading / packing this entry
            *out = e.dev_raw;
            *entry_idx = i;
            return QNN_OK;
        }
    }
    size_t slot = wc.entries.size();
    if (slot >= kMaxResident) {  // recycle the least recently used entry
        slot = 0;
        for (size_t i = 1; i < wc.entries.size(); ++i)
            if (wc.entries[i].tick < wc.entries[slot].tick) slot = i;
        ResidentWeights& old = wc.entries[slot];
        if (old.last_use) cudaStreamWaitEvent(st, old.last_use, 0);
        if (old.dev_raw) cudaFreeAsync(old.dev_raw, st);
        if (old.dev_packed) cudaFreeAsync(old.dev_packed, st);
        cudaEvent_t ev = old.last_use;
        old = ResidentWeights();
        old.last_use = ev;
    } else {
        wc.entries.emplace_back();
    }
    ResidentWeights& e = wc.entries[slot];
    if (!e.last_use && cudaEventCreateWithFlags(&e.last_use, cudaEventDisableTiming) != cudaSuccess) {
        set_error("event creation failed");
        return QNN_E_CUDA;
    }
    int rc = stream_scratch_alloc(reinterpret_cast<void**>(&e.dev_raw), bytes, st);
    if (rc) return rc;
    cudaError_t ce = cudaMemcpyAsync(e.dev_raw, host, bytes, cudaMemcpyHostToDevice, st);
    if (ce != cudaSuccess) {
        set_error("weight upload failed: %s", cudaGetErrorString(ce));
        return QNN_E_CUDA;
    }
    cudaEventRecord(e.last_use, st);
    e.host = host;
    e.bytes = bytes;
    e.hash = h;
    e.tick = ++wc.tick;
    *out = e.dev_raw;
    *entry_idx = slot;
    return QNN_OK;
}

// the packed image kept next to a resident kernel: (re)built on `st` when the signature (kernel family, tile, math) differs
int resident_packed(int dev, size_t entry_idx, uint64_t sig, size_t bytes, cudaStream_t st, void** out, bool* fresh) {
    WeightCache& wc = g_weights[dev];
    std::lock_guard<std::mutex> lock(wc.mu);
    ResidentWeights& e = wc.entries[entry_idx];
    *fresh = false;
    if (e.dev_packed && e.pack_sig == sig) {
        *out = e.dev_packed;
        return QNN_OK;
    }
    if (e.dev_packed) {
        if (e.last_use) cudaStreamWaitEvent(st, e.last_use, 0);
        cudaFreeAsync(e.dev_packed, st);
        e.dev_packed = nullptr;
    }
    int rc = stream_scratch_alloc(&e.dev_packed, bytes, st);
    if (rc) return rc;
    e.pack_sig = sig;
    *out = e.dev_packed;
    *fresh = true;
    return QNN_OK;
}

void mark_used(int dev, size_t entry_idx, cudaStream_t st) {
    WeightCache& wc = g_weights[dev];
    std::lock_guard<std::mutex> lock(wc.mu);
    if (entry_idx < wc.entries.size() && wc.entries[entry_idx].last_use) cudaEventRecord(wc.entries[entry_idx].last_use, st);
}

int forward_host(const Geom& g, int rank, int math, int algo, size_t nx, size_t nw, size_t nb, size_t ny,
                 const float* xh, const float* wh, const float* bh, float* yh, cudaStream_t st) {
    if (!valid_math(math) || !valid_algo(algo)) {
        set_error("unknown math mode %d / algo %d", math, algo);
        return QNN_E_INVALID;
    }
    if (ny == 0) return QNN_OK;
    if (!xh || !wh || !yh) {
        set_error("host x, kernel and y must not be NULL");
        return QNN_E_INVALID;
    }
    const int dev = current_device();
    // resident weights (+ packed image when a tensor-core kernel takes the problem)
    float *wd = nullptr, *bd = nullptr;
    size_t w_idx = 0, b_idx = 0;
    int rc = resident_weights(dev, wh, nw, st, &wd, &w_idx);
    if (!rc && bh) rc = resident_weights(dev, bh, nb, st, &bd, &b_idx);
    if (rc) return rc;
    const int x3 = math == QNN_MATH_3XTF32;
    const int kern = forward_kernel(g, rank, math, algo);
    void* packed = nullptr;
    if (kern == kKernTc || kern == kKernTc2d) {
        const size_t pbytes = kern == kKernTc ? tc_packed_bytes(g, rank, x3) : tc2d_packed_bytes(g, rank, x3);
        bool fresh = false;
        rc = resident_packed(dev, w_idx, ((uint64_t)kern << 8) | (uint64_t)x3 | ((uint64_t)pbytes << 16), pbytes, st, &packed, &fresh);
        if (!rc && fresh) {
            rc = kern == kKernTc ? tc_pack(g, rank, x3, 0, wd, packed, st) : tc2d_pack(g, rank, x3, 0, wd, packed, st);
            mark_used(dev, w_idx, st);  // later users of the entry (any stream) wait for the image
        }
        if (rc) return rc;
    }
    float *xd = nullptr, *yd = nullptr;
    if ((rc = stream_scratch_alloc(reinterpret_cast<void**>(&xd), align256(nx * 4), st))) return rc;
    if ((rc = stream_scratch_alloc(reinterpret_cast<void**>(&yd), align256(ny * 4), st))) {
        cudaFreeAsync(xd, st);
        return rc;
    }
    // units that can be cut independently: samples (conv) or rows (dense: batch == 1, rows live in in_sp[2])
    const bool dense = g.conj_w && g.batch == 1 && g.k[2] == 1;
    const long long units = dense ? g.in_sp[2] : g.batch;
    const size_t x_unit = nx / (size_t)units, y_unit = ny / (size_t)units;
    int chunks = 1;
    HostPipe* pipe = nullptr;
    if ((nx + ny) * 4 >= (size_t)8 << 20 && units >= 8 && (pipe = borrow_pipe(dev))) {
        chunks = units >= 64 ? 8 : 4;  // measured on cfg 2: 1 / 2 / 4 / 8 / 16 chunks -> 2.04 / 1.70 / 1.59 / 1.54 / 1.65 ms
                                       // (67 MB of y over PCIe is ~1.45 ms on its own; uneven cuts gained nothing)
        if (const char* env = getenv("QNN_HOST_CHUNKS")) {  // tuning knob: 1..16 pipeline chunks
            const int v = atoi(env);
            if (v >= 1 && v <= 16 && v <= units) chunks = v;
        }
    }
    cudaError_t e = cudaSuccess;
    auto finish = [&](int code) {
        mark_used(dev, w_idx, st);
        if (bh) mark_used(dev, b_idx, st);
        cudaFreeAsync(xd, st);
        cudaFreeAsync(yd, st);
        if (pipe) return_pipe(dev, pipe);
        if (code == QNN_OK && e != cudaSuccess) {
            set_error("host staging failed: %s", cudaGetErrorString(e));
            return (int)QNN_E_CUDA;
        }
        return code;
    };
    if (chunks == 1 || !pipe) {
        if ((e = cudaMemcpyAsync(xd, xh, nx * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return finish(QNN_OK);
        rc = run_forward(g, rank, math, algo, xd, wd, packed, bd, yd, st);
        if (rc) return finish(rc);
        if ((e = cudaMemcpyAsync(yh, yd, ny * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return finish(QNN_OK);
        e = cudaStreamSynchronize(st);
        return finish(QNN_OK);
    }
    // the copy streams must not run ahead of work already queued on the caller's stream (scratch allocation, uploads)
    if ((e = cudaEventRecord(pipe->ev_start, st)) != cudaSuccess) return finish(QNN_OK);
    if ((e = cudaStreamWaitEvent(pipe->in, pipe->ev_start, 0)) != cudaSuccess) return finish(QNN_OK);
    if ((e = cudaStreamWaitEvent(pipe->out, pipe->ev_start, 0)) != cudaSuccess) return finish(QNN_OK);
    for (int c = 0; c < chunks; ++c) {
        const long long u0 = units * c / chunks, u1 = units * (c + 1) / chunks;
        if (u1 == u0) continue;
        Geom gc = g;
        if (dense)
            gc.in_sp[2] = gc.out_sp[2] = (int)(u1 - u0);
        else
            gc.batch = (int)(u1 - u0);
        if ((e = cudaMemcpyAsync(xd + u0 * x_unit, xh + u0 * x_unit, (size_t)(u1 - u0) * x_unit * 4,
                                 cudaMemcpyHostToDevice, pipe->in)) != cudaSuccess) return finish(QNN_OK);
        if ((e = cudaEventRecord(pipe->ev_in[c], pipe->in)) != cudaSuccess) return finish(QNN_OK);
        if ((e = cudaStreamWaitEvent(st, pipe->ev_in[c], 0)) != cudaSuccess) return finish(QNN_OK);
        rc = run_forward(gc, rank, math, algo, xd + u0 * x_unit, wd, packed, bd, yd + u0 * y_unit, st);
        if (rc) return finish(rc);
        if ((e = cudaEventRecord(pipe->ev_k[c], st)) != cudaSuccess) return finish(QNN_OK);
        if ((e = cudaStreamWaitEvent(pipe->out, pipe->ev_k[c], 0)) != cudaSuccess) return finish(QNN_OK);
        if ((e = cudaMemcpyAsync(yh + u0 * y_unit, yd + u0 * y_unit, (size_t)(u1 - u0) * y_unit * 4,
                                 cudaMemcpyDeviceToHost, pipe->out)) != cudaSuccess) return finish(QNN_OK);
    }
    // the caller's stream joins the copy-out stream (the stream-ordered frees below must come after the last D2H)
    if ((e = cudaEventRecord(pipe->ev_out, pipe->out)) != cudaSuccess) return finish(QNN_OK);
    if ((e = cudaStreamWaitEvent(st, pipe->ev_out, 0)) != cudaSuccess) return finish(QNN_OK);
    e = cudaStreamSynchronize(st);
    return finish(QNN_OK);
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
    if (!valid_math(d->math) || !valid_algo(d->algo) || empty_out(g)) return 0;
    const int kern = forward_kernel(g, d->rank, d->math, d->algo);
    return kern == kKernTc || kern == kKernTc2d;
}

int qnn_conv_forward_kernel(const qnn_conv_desc* d) {
    Geom g;
    if (build_geom(d, &g)) return QNN_E_INVALID;
    if (!valid_math(d->math) || !valid_algo(d->algo)) return QNN_E_INVALID;
    return empty_out(g) ? QNN_KERNEL_GENERAL : forward_kernel(g, d->rank, d->math, d->algo);
}

int qnn_conv_work_split(const qnn_conv_desc* d, int32_t* out4) {
    Geom g;
    if (build_geom(d, &g)) return QNN_E_INVALID;
    if (!valid_math(d->math) || !valid_algo(d->algo) || !out4) return QNN_E_INVALID;
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    if (empty_out(g)) return QNN_KERNEL_GENERAL;
    const int kern = forward_kernel(g, d->rank, d->math, d->algo);
    const int x3 = d->math == QNN_MATH_3XTF32;
    long long items = 0;
    int f_tile = 0;
    if (kern == kKernTc) {
        items = tc_work_items(g);
        f_tile = tc_plan(g, d->rank, x3).f_tile;
    } else if (kern == kKernTc2d) {
        const Tc2dPlan pl = tc2d_plan(g, d->rank, x3);
        items = tc2d_work_items(g, pl);
        f_tile = pl.f_tile;
    } else {
        return kern;
    }
    if (items > 0x7fffffffLL / 64) return QNN_E_UNSUPPORTED;
    const WorkSplit ws = plan_work_split((int)items, f_tile, num_sms());
    out4[0] = ws.grid;
    out4[1] = ws.full_rounds;
    out4[2] = ws.rem;
    out4[3] = ws.split;
    return kern;
}

int qnn_dense_uses_tensor_cores(int64_t rows, int32_t in_q, int32_t q_units) {
    Geom g;
    if (build_dense_geom(rows, in_q, q_units, QNN_ACT_LINEAR, &g)) return 0;
    return tc_plan(g, 1, 0).ok;
}

int qnn_dense_forward_kernel(int64_t rows, int32_t in_q, int32_t q_units, int32_t activation, int32_t math, int32_t algo) {
    Geom g;
    if (build_dense_geom(rows, in_q, q_units, activation, &g)) return QNN_E_INVALID;
    if (!valid_math(math) || !valid_algo(algo)) return QNN_E_INVALID;
    return empty_out(g) ? QNN_KERNEL_GENERAL : forward_kernel(g, 1, math, algo);
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
    *dx_tc = a != kKernGeneral;
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
    *dx_tc = a != kKernGeneral;
    *dkernel_tc = b;
    return QNN_OK;
}

int qnn_conv_forward(const qnn_conv_desc* d, const float* x, const float* kernel, const float* bias, float* y,
                     void* stream) {
    Geom g;
    int rc = build_geom(d, &g);
    if (rc) return rc;
    return run_forward(g, d->rank, d->math, d->algo, x, kernel, nullptr, bias, y, static_cast<cudaStream_t>(stream));
}

int qnn_dense_forward(int64_t rows, int32_t in_q, int32_t q_units, const float* x, const float* kernel,
                      const float* bias, int32_t activation, int32_t math, int32_t algo, float* y, void* stream) {
    Geom g;
    int rc = build_dense_geom(rows, in_q, q_units, activation, &g);
    if (rc) return rc;
    return run_forward(g, 1, math, algo, x, kernel, nullptr, bias, y, static_cast<cudaStream_t>(stream));
}

// ---- packed kernel images (caller-owned cache)
namespace {
// geometry + kernel family of the problem a packed image of `kind` feeds: the layer's own forward, or the transposed
// problem of its data gradient
int packed_problem(const Geom& g, int rank, int math, int algo, int kind, Geom* gp, int* kern) {
    if (kind != QNN_PACK_FORWARD && kind != QNN_PACK_DGRAD) {
        set_error("unknown packed-kernel kind %d", kind);
        return QNN_E_INVALID;
    }
    if (!valid_math(math) || !valid_algo(algo)) {
        set_error("unknown math mode %d / algo %d", math, algo);
        return QNN_E_INVALID;
    }
    *kern = kKernGeneral;
    *gp = g;
    if (empty_out(g)) return QNN_OK;
    if (kind == QNN_PACK_FORWARD) {
        *kern = forward_kernel(g, rank, math, algo);
        if (*kern == kKernSmallK) *kern = kKernGeneral;  // no packed form
    } else {
        int dxk = 0, dwk = 0;
        backward_selection(g, rank, math, algo, &dxk, &dwk);
        *kern = dxk;
        *gp = transposed_geom(g);
    }
    return QNN_OK;
}
size_t packed_bytes_of(const Geom& g, int rank, int math, int algo, int kind) {
    Geom gp;
    int kern;
    if (packed_problem(g, rank, math, algo, kind, &gp, &kern) || kern == kKernGeneral) return 0;
    const int x3 = math == QNN_MATH_3XTF32;
    return kern == kKernTc ? tc_packed_bytes(gp, rank, x3) : tc2d_packed_bytes(gp, rank, x3);
}
int pack_of(const Geom& g, int rank, int math, int algo, int kind, const float* kernel, void* packed, cudaStream_t st) {
    Geom gp;
    int kern;
    int rc = packed_problem(g, rank, math, algo, kind, &gp, &kern);
    if (rc) return rc;
    if (kern == kKernGeneral) {
        set_error("this problem has no packed kernel image (it runs on the general kernel): qnn_*_packed_bytes returned 0");
        return QNN_E_UNSUPPORTED;
    }
    if (!kernel || !packed) {
        set_error("kernel and packed must not be NULL");
        return QNN_E_INVALID;
    }
    const int x3 = math == QNN_MATH_3XTF32, tr = kind == QNN_PACK_DGRAD;
    return kern == kKernTc ? tc_pack(gp, rank, x3, tr, kernel, packed, st) : tc2d_pack(gp, rank, x3, tr, kernel, packed, st);
}
}  // namespace

size_t qnn_conv_packed_bytes(const qnn_conv_desc* d, int32_t kind) {
    Geom g;
    if (build_geom(d, &g)) return 0;
    return packed_bytes_of(g, d->rank, d->math, d->algo, kind);
}

int qnn_conv_pack(const qnn_conv_desc* d, int32_t kind, const float* kernel, void* packed, void* stream) {
    Geom g;
    int rc = build_geom(d, &g);
    if (rc) return rc;
    return pack_of(g, d->rank, d->math, d->algo, kind, kernel, packed, static_cast<cudaStream_t>(stream));
}

int qnn_conv_forward_packed(const qnn_conv_desc* d, const float* x, const float* kernel, const void* packed,
                            const float* bias, float* y, void* stream) {
    Geom g;
    int rc = build_geom(d, &g);
    if (rc) return rc;
    return run_forward(g, d->rank, d->math, d->algo, x, kernel, packed, bias, y, static_cast<cudaStream_t>(stream));
}

size_t qnn_dense_packed_bytes(int64_t rows, int32_t in_q, int32_t q_units, int32_t math, int32_t algo, int32_t kind) {
    Geom g;
    if (build_dense_geom(rows, in_q, q_units, QNN_ACT_LINEAR, &g)) return 0;
    return packed_bytes_of(g, 1, math, algo, kind);
}

int qnn_dense_pack(int64_t rows, int32_t in_q, int32_t q_units, int32_t math, int32_t algo, int32_t kind,
                   const float* kernel, void* packed, void* stream) {
    Geom g;
    int rc = build_dense_geom(rows, in_q, q_units, QNN_ACT_LINEAR, &g);
    if (rc) return rc;
    return pack_of(g, 1, math, algo, kind, kernel, packed, static_cast<cudaStream_t>(stream));
}

int qnn_dense_forward_packed(int64_t rows, int32_t in_q, int32_t q_units, const float* x, const float* kernel,
                             const void* packed, const float* bias, int32_t activation, int32_t math, int32_t algo,
                             float* y, void* stream) {
    Geom g;
    int rc = build_dense_geom(rows, in_q, q_units, activation, &g);
    if (rc) return rc;
    return run_forward(g, 1, math, algo, x, kernel, packed, bias, y, static_cast<cudaStream_t>(stream));
}

namespace {
int backward_checks(const Geom& g, const float* x, const float* kernel, const float* y, const float* dy, float* dx,
                    float* dkernel, float* dbias, bool* nothing) {
    *nothing = false;
    if (g.act != QNN_ACT_LINEAR && g.act != QNN_ACT_RELU) {
        set_error("backward supports linear and relu activations only");
        return QNN_E_UNSUPPORTED;
    }
    if (!x || !kernel || !y || !dy) {
        if (empty_out(g) && !dkernel && !dbias && !dx) {
            *nothing = true;
            return QNN_OK;
        }
        set_error("x, kernel, y and dy must not be NULL");
        return QNN_E_INVALID;
    }
    return QNN_OK;
}
}  // namespace

int qnn_conv_backward(const qnn_conv_desc* d, const float* x, const float* kernel, const float* y, const float* dy,
                      float* dx, float* dkernel, float* dbias, void* stream) {
    return qnn_conv_backward_packed(d, x, kernel, nullptr, y, dy, dx, dkernel, dbias, stream);
}

int qnn_conv_backward_packed(const qnn_conv_desc* d, const float* x, const float* kernel, const void* packed_dgrad,
                             const float* y, const float* dy, float* dx, float* dkernel, float* dbias, void* stream) {
    Geom g;
    int rc = build_geom(d, &g);
    if (rc) return rc;
    bool nothing;
    if ((rc = backward_checks(g, x, kernel, y, dy, dx, dkernel, dbias, &nothing)) || nothing) return rc;
    return run_backward(g, d->rank, d->math, d->algo, x, kernel, packed_dgrad, y, dy, dx, dkernel, dbias,
                        static_cast<cudaStream_t>(stream));
}

int qnn_dense_backward(int64_t rows, int32_t in_q, int32_t q_units, const float* x, const float* kernel,
                       const float* y, const float* dy, int32_t activation, int32_t math, int32_t algo, float* dx,
                       float* dkernel, float* dbias, void* stream) {
    return qnn_dense_backward_packed(rows, in_q, q_units, x, kernel, nullptr, y, dy, activation, math, algo, dx, dkernel,
                                     dbias, stream);
}

int qnn_dense_backward_packed(int64_t rows, int32_t in_q, int32_t q_units, const float* x, const float* kernel,
                              const void* packed_dgrad, const float* y, const float* dy, int32_t activation,
                              int32_t math, int32_t algo, float* dx, float* dkernel, float* dbias, void* stream) {
    Geom g;
    int rc = build_dense_geom(rows, in_q, q_units, activation, &g);
    if (rc) return rc;
    bool nothing;
    if ((rc = backward_checks(g, x, kernel, y, dy, dx, dkernel, dbias, &nothing)) || nothing) return rc;
    return run_backward(g, 1, math, algo, x, kernel, packed_dgrad, y, dy, dx, dkernel, dbias,
                        static_cast<cudaStream_t>(stream));
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
