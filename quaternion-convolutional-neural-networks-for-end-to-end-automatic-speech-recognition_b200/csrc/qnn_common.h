// Internal declarations shared by the kernels and the C-ABI translation unit.
#pragma once
#include <cstdlib>
#include <cstdint>
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include <cuda_runtime.h>
#include "../../include/qnn.h"

namespace qnn {

// Geometry of one quaternion convolution, normalised to three spatial axes (leading axes have extent 1 for rank < 3).
// A dense layer is the degenerate case: one position per row, one tap, `conj_w` set (transposed Hamilton table).
struct Geom {
    int batch;
    int in_sp[3], out_sp[3], k[3], s[3], d[3], pad_lo[3];
    int in_q, F;
    int channels_first;
    int act;
    int conj_w;  // 1 = dense convention y = conj(W) (x) x  (complexnn/dense.py:139-143), 0 = conv  y = W (x) x
    int dense;   // 1 = a QuaternionDense problem (or its transposed data-gradient problem): kernel selection for it must not
                 // depend on the activation, which the packed-image queries of the dense ABI do not carry
};

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
// stream-ordered device scratch (cudaMallocAsync on the default pool, which is told once to keep freed blocks);
// release with cudaFreeAsync on the same stream.  Returns QNN_OK or QNN_E_CUDA (error text set).
int stream_scratch_alloc(void** ptr, size_t bytes, cudaStream_t st);
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
// Per-DEVICE state (one process may drive several GPUs): SM count of the current device, and the opt-in to more than
// 48 KB of dynamic shared memory, recorded once per (device, kernel).  ensure_dynamic_smem returns QNN_OK / QNN_E_CUDA.
constexpr int kMaxDevices = 64;
int current_device();
int num_sms();
int ensure_dynamic_smem(const void* kernel, int bytes);

// general (CUDA-core, fp32) kernels: any rank / stride / dilation / padding / data_format
int general_forward(const Geom& g, const float* x, const float* w, const float* bias, float* y, cudaStream_t st);
int general_backward(const Geom& g, const float* x, const float* w, const float* y, const float* dy, float* dx, float* dw,
                     float* db, cudaStream_t st);

// small-K forward (CUDA cores, warp-shuffle tap reuse): channels_last rank 1 / dense, stride 1, in_q < 4 (qnn_smallk.cu)
struct SmallKPlan {
    int ok;
    int pc;  // output positions per warp run
    size_t smem_bytes;
    const char* why;
};
SmallKPlan smallk_plan(const Geom& g, int rank);
int smallk_forward(const Geom& g, int rank, const float* x, const float* w, const float* bias, float* y, cudaStream_t st);

// helper of the tensor-core backward path: dz = dy * act'(y) with the bias gradient folded in (either output may be
// NULL); channels_last: rows x C, channels_first: [n][C][S] with S positions per channel
int dz_bgrad(const float* y, const float* dy, float* dz, float* db, long long rows, int C, int relu, cudaStream_t st);
int dz_bgrad_cf(const float* y, const float* dy, float* dz, float* db, int batch, int C, long long S, int relu,
                cudaStream_t st);
// [batch][C][S] -> [batch][S][C]: channels_last scratch copies of channels_first tensors for the kernel gradient
int cf_to_cl(const float* in, float* out, int batch, int C, long long S, cudaStream_t st);

// tensor-core kernel (tcgen05 / TMEM / TMA): channels_last rows with taps along the innermost spatial axis.
// `x3` selects 3xTF32 arithmetic (hi / lo operand split, three MMAs per block) instead of plain TF32.
// How a persistent tensor-core forward kernel spreads its work items (tiles x filter tiles) over the SMs: `full_rounds`
// whole items per CTA (item = cta + k * grid), then `rem` items in a last round; `split`: that round's items are split along
// the filters into 2 * rem half-width items (WorkItem in qnn_hamilton_tc.cu, Work in qnn_hamilton_tc2d.cu) -- only 64-wide
// filter tiles split, and only when the halves keep no more CTAs busy than there are SMs.  QNN_TC_NOSPLIT=1 switches it off.
struct WorkSplit {
    int grid, full_rounds, rem, split;
};
inline WorkSplit plan_work_split(int n_items, int f_tile, int sms) {
    static const bool no_split = [] { const char* e = getenv("QNN_TC_NOSPLIT"); return e && atoi(e) != 0; }();
    WorkSplit w{};
    if (n_items >= sms) {
        w.grid = sms;
        w.full_rounds = n_items / sms;
        w.rem = n_items % sms;
        w.split = (!no_split && f_tile == 64 && w.rem > 0 && 2 * w.rem <= sms) ? 1 : 0;
    } else {
        w.full_rounds = 0;
        w.rem = n_items;
        w.split = (!no_split && f_tile == 64 && 2 * w.rem <= sms) ? 1 : 0;
        w.grid = w.split ? 2 * w.rem : w.rem;
    }
    return w;
}

struct TcPlan {
    int ok;           // shape qualifies
    int f_tile;       // filters per pass (<= 64)
    int n_ftiles;
    int in_q_pad;     // in_q rounded up to 8
    int pad_x;        // (unused since round 2: ragged channel counts are read in place)
    int ragged;       // in_q % 4 != 0: un-swizzled 36-channel boxes of the flat row, 4-byte converter loads
    int rows_in;      // input rows per tile: 127 * stride + (taps-1) * dilation + 1 (rounded up to whole boxes)
    int box_rows;     // rows per TMA box (<= 256)
    int x_stages;
    int n_st;         // output staging tiles (2, 4 or 8)
    size_t smem_bytes;
    size_t w_bytes;       // one part (hi or lo) of one filter tile's packed image
    size_t packed_bytes;  // whole packed kernel image (all filter tiles, hi [+ lo])
    const char* why;  // reason when !ok
};
TcPlan tc_plan(const Geom& g, int rank, int x3);
long long tc_work_items(const Geom& g);  // work items per filter-tile pass (tiles of 128 output positions)
void tc_set_trace(void* device_buffer, size_t bytes);
// x[rows][4][in_q] -> xp[rows][4][xq], xq = in_q rounded up to 4, new channels zero (ragged channel counts)
int pad_x_channels(const float* x, float* xp, long long rows, int in_q, int xq, cudaStream_t st);
// packed kernel image: K-major tf32 core matrices per filter tile (see qnn_hamilton_tc.cu); `transposed` reads the stored
// kernel of the layer whose data gradient `g` describes
size_t tc_packed_bytes(const Geom& g, int rank, int x3);
int tc_pack(const Geom& g, int rank, int x3, int transposed, const float* w, void* packed, cudaStream_t st);
int tc_forward_packed(const Geom& g, int rank, int x3, const float* x, const void* packed, const float* bias, float* y,
                      cudaStream_t st);
int tc_forward(const Geom& g, int rank, int x3, int transposed, const float* x, const float* w, const float* bias, float* y,
               cudaStream_t st);

// tensor-core kernel gradient (channels_last rank 1 / dense, stride 1): contraction over positions, four persistent
// accumulators D_c = dL/df_c in tensor memory, vector reductions into dW
struct WgradPlan {
    int ok;
    int f_tile, n_ftiles, n_mblk;
    int rows;  // x stage rows: 32 + (taps - 1) * dilation
    int pad_x;  // in_q % 4 != 0 and 4 in_q > 256: x goes through the channel-padding pre-pass
    int flat_x; // in_q % 4 != 0 and 4 in_q <= 256: the x stage is the flat channel row (no pre-pass)
    int x_stages;
    size_t x_stage_bytes, smem_bytes;
    const char* why;
};
WgradPlan wgrad_plan(const Geom& g, int rank, int x3);
void wgrad_set_trace(void* device_buffer, size_t bytes);
int wgrad_tc(const Geom& g, int rank, int x3, const float* x, const float* dz, float* dw, cudaStream_t st,
             const float* yfwd = nullptr, float* dz_out = nullptr, float* db = nullptr);

// tensor-core kernel for channels_first tensors (rank 1 / 2, stride 1): streamed sub-filters, transposing converters
struct Tc2dPlan {
    int ok;
    int f_tile, n_ftiles;
    int wbox;        // x box per channel: one row of 128 + halo positions (rounded up to 4)
    int xshift;      // columns the box starts left of the first tap (16-byte alignment of the TMA start)
    int pad_rows;    // channels_first row length % 4 != 0: x / y go through row-padded scratch copies
    int pad_q;       // in_q % 8 != 0: x goes through a quaternion-channel padding pre-pass (zero channels)
    int rag;         // channels_last, in_q % 8 != 0, in_q >= 8: read in place through twelve-channel boxes (no pre-pass)
    int x_stages;
    size_t x_stage_bytes, smem_bytes;
    size_t packed_bytes;  // packed kernel image (hi [+ lo] blocks per (filter tile, 8-channel chunk, tap))
    const char* why;
};
Tc2dPlan tc2d_plan(const Geom& g, int rank, int x3);
long long tc2d_work_items(const Geom& g, const Tc2dPlan& pl);  // (tile, filter tile) items
void tc2d_set_trace(void* device_buffer, size_t bytes);
size_t tc2d_packed_bytes(const Geom& g, int rank, int x3);
int tc2d_pack(const Geom& g, int rank, int x3, int transposed, const float* w, void* packed, cudaStream_t st);
int tc2d_forward_packed(const Geom& g, int rank, int x3, const float* x, const void* packed, const float* bias, float* y,
                        cudaStream_t st);
int tc2d_forward(const Geom& g, int rank, int x3, int transposed, const float* x, const float* w, const float* bias, float* y,
                 cudaStream_t st);

}  // namespace qnn

// ---------------------------------------------------------------------------------------------------------------------
// activation (device): keras.activations semantics
// ---------------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace qnn {
__device__ __forceinline__ float act_apply(float v, int act) {
    switch (act) {
        case QNN_ACT_RELU: return fmaxf(v, 0.f);
        case QNN_ACT_TANH: return tanhf(v);
        case QNN_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case QNN_ACT_HARD_SIGMOID: return fminf(fmaxf(0.2f * v + 0.5f, 0.f), 1.f);
        case QNN_ACT_SOFTPLUS: return fmaxf(v, 0.f) + log1pf(expf(-fabsf(v)));
        case QNN_ACT_SOFTSIGN: return v / (1.f + fabsf(v));
        case QNN_ACT_ELU: return v > 0.f ? v : expm1f(v);
        case QNN_ACT_SELU: return 1.0507009873554805f * (v > 0.f ? v : 1.6732632423543772f * expm1f(v));
        case QNN_ACT_EXPONENTIAL: return expf(v);
        default: return v;
    }
}
}  // namespace qnn
#endif
