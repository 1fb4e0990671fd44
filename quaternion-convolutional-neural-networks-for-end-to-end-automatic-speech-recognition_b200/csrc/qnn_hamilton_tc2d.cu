// Fused Hamilton convolution for channels_first tensors on the 5th-generation tensor cores (sm_100a only):
// QuaternionConv2D (and Conv1D) with data_format='channels_first', stride 1 -- BASELINE configs[4],
// models/interspeech_model.py:60-61,97 in the reference.
//
// Problem: x[nb, 4*in_q, H, W], stored un-expanded kernel w[KH, KW, in_q, 4F], y[nb, 4F, Ho, Wo] =
// act(bias + sum_tap sum_a sum_b +-x_a(h + kh*dh - ph, w + kw*dw - pw) . f_{a^b}[tap]).  As in qnn_hamilton_tc.cu the
// 4in_q x 4F real weight of complexnn/conv.py:327-331 never exists: 16 signed tcgen05.mma per k-step share four A
// operands (input components) and four B operands (sub-filters), sign = negate-B bit of the instruction descriptor.
//
// What differs from the channels_last kernel:
//   * positions are the CONTIGUOUS axis of x and y.  One tile = 128 consecutive output columns of one output row.
//     An x stage is a 4-D TMA box (W: 128 + halo, one input row = one kernel row kh, 8 quaternion channels, the 4
//     components of a sample) landing as [channel][w]; a converter thread owns one position and walks the channels (conflict-free
//     4-byte shared loads), so the transpose to "position = TMEM lane, channel = TMEM column" costs nothing extra.
//     Rows / columns outside the image are TMA zero fill = the convolution's zero padding.
//   * a whole layer's sub-filters (cfg 5: 1.18 MB) do not fit in shared memory, so they are STREAMED: a tiny
//     pre-pass (k_pack_w2d) re-orders the stored kernel into the K-major core-matrix image the MMA wants, rounded
//     to nearest tf32, one 8 KB block per (filter tile, 8-channel chunk, tap); the main kernel pulls a block with one
//     cp.async.bulk into the B slot that pairs with the A slot of the same (chunk, tap).  A and B rings share their
//     index and their "empty" barrier (the same 16 MMAs consume both).
//   * the k loop is chunk-major: a stage holds all FOUR components of 8 quaternion channels, so one 8 KB sub-filter
//     block feeds 16 MMAs (every (a, b) pair) and is read from L2 exactly once per tile.
//   * the epilogue transposes back through a [32 channels][128 positions] staging tile and TMA-stores it (4-D box),
//     which also clips ragged tiles.
// Work item = (tile, filter tile of <= 64 filters); persistent CTAs stride over the items.
//
// Template parameter CL selects the channels_last rank-2 variant of the same main loop -- the x stage is four 5-D boxes
// (8 q, one component, 128 + halo columns, 1 row, 1 sample), dense [component][w][8 q] (no swizzle; read with 16-byte
// loads whose order alternates between lanes), the epilogue stages [128 positions][32 channels] rows.
// Warp roles and the TMEM plan are those of qnn_hamilton_tc.cu: warps 0-15 epilogue, 16-19 MMA issuers (one per output
// component), 20-27 converters (two groups on alternate stages), 28 / 29 producers (x stages / sub-filter blocks); TMEM [0,256) accumulators, [256,512) eight A slots.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include "qnn_common.h"
#include "qnn_ptx.cuh"
#include "qnn_tmap.h"

namespace qnn {
namespace {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kThreads = 1024;
constexpr int kEpiThreads = 512;
constexpr int kMaxSlots = 8;   // TF32: eight 32-column A slots (and B blocks); 3xTF32: four 64-column ones (hi | lo)
constexpr int kAccCols = 256;
constexpr int kMaxXStages = 6;
constexpr int kMaxTapBatch = 4;
constexpr int kStagingBytes = 32 * kTileM * 4;  // [32 channels][128 positions] fp32
constexpr uint32_t kSmemLimit = 232448;
// register budget: 1024 x 64 at launch = 2 x 128 x kRegsWg0 (issuers, producers) + 256 x kRegsWg1 (converters) + 512 x kRegsEpi
constexpr int kRegsWg0 = 24, kRegsWg1 = 40, kRegsEpi = 96;
static_assert(256 * kRegsWg0 + 256 * kRegsWg1 + 512 * kRegsEpi <= 1024 * 64, "register pool");

constexpr uint32_t kNegConv = (1u << 4) | (1u << 7) | (1u << 8) | (1u << 9) | (1u << 12) | (1u << 14);
constexpr uint32_t transpose_bits(uint32_t m) {
    uint32_t t = 0;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
            if ((m >> (a * 4 + b)) & 1u) t |= 1u << (b * 4 + a);
    return t;
}
constexpr uint32_t kNegDense = transpose_bits(kNegConv);

enum { kActLinear = 0, kActRelu = 1, kActGeneric = 2 };
enum { kWarpAlloc = 16, kWarpIssuer0 = 16, kWarpConv0 = 20, kWarpProd0 = 28 /* 28: x TMA, 29: sub-filter blocks */ };

// Optional per-CTA event trace (qnn_debug_trace, tools/tc2d_trace.py): clock64() slots per CTA
constexpr int kTraceSlots = 256;
enum { kTrStart = 0, kTrSetup = 1, kTrEnd = 2, kTrItem = 8 /* + 4*item: mma acc_empty passed, mma committed, epilogue got
       acc, epilogue stored */, kTrIssue = 64 /* + 2*slot of the CTA's second item: a_full passed, b_full passed */,
       kTrConv = 160 /* + 4*k, k-th own stage of converter group 0 in the second item: x_full passed, a_empty passed,
                        stores issued, slot released */ };

struct P2 {
    unsigned long long* trace;
    int n_items, n_ftiles;
    int tiles_w, Ho, Do;   // Do = 1 below rank 3
    int KH, KW, taps, dh, dw, pad_h, pad_w;
    int KD, dd, pad_d, rank3;  // rank 3 (channels_first only): a third, outermost spatial axis -- 5-D tensor maps
    int n_qc;      // 8-channel chunks per item
    int n_stages;  // x stages per item = n_qc * KH: one kernel row of one chunk each
    int F, f_tile;
    int wbox;        // x box: one row of wbox positions per channel
    int xshift;      // the box starts xshift (0..3) columns left of the first tap: its start must be 16-byte aligned
    int x_stages, x_stage_bytes;
    int act, has_bias;
    int comp_stride;       // channels_last: bytes between the component blocks of an x stage (wbox * 32 rounded up to 128)
    int handshake;         // KW >= A slots: converter groups hand over stage by stage (see the converter role)
    uint32_t b_blk_bytes;  // one B slot: 4 sub-filters x 2 k-groups x f_tile x 16 B (3xTF32: twice that, hi block | lo block)
    int rag;               // channels_last, in_q % 8 != 0, read in place: a component's 8 channels of a stage do not start on a
                           // 16-byte boundary -- the stage is four un-swizzled boxes of TWELVE flat channels (48 bytes per
                           // position) starting at the aligned channel at or below, the converter selects its 8 by the shift
    int in_q;              // the tensor's own quaternion channel count (n_qc counts the zero-padded one)
    int full_rounds, rem;  // work of a CTA: full_rounds items (item = cta + k * grid), then `rem` items in a last round
    int split;             // the last round's items are split along the filters, see Work
};

struct __align__(8) Bars {
    uint64_t x_full[kMaxXStages], x_empty[kMaxXStages];
    uint64_t a_full[kMaxSlots], b_full[kMaxSlots], a_empty[kMaxSlots];
    uint64_t acc_full, acc_empty;
    uint32_t tmem_base;
};

__device__ __forceinline__ void trace(const P2& p, int slot) {
    if (p.trace && slot < kTraceSlots) p.trace[(size_t)blockIdx.x * kTraceSlots + slot] = (unsigned long long)clock64();
}

template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc,
                                       uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 d;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 d, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], d, %4, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// plain (1-D) bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// the run-time activations (tanh, sigmoid, elu, ...) as ONE out-of-line function: inlined 64 times per epilogue thread they
// made each generic-activation kernel 31 k instructions and dominated the library's compile time
__device__ __noinline__ float act_apply_call(float v, int act) { return act_apply(v, act); }

template <int ACT>
__device__ __forceinline__ float activate(float v, int act_rt) {
    if (ACT == kActLinear) return v;
    if (ACT == kActRelu) return fmaxf(v, 0.f);
    return act_apply_call(v, act_rt);
}

__device__ __forceinline__ void split_tf32(uint32_t v, uint32_t& hi, uint32_t& lo) {
    hi = (v + 0x1000u) & 0xffffe000u;                                                   // rn_tf32(v), low bits cleared
    lo = __float_as_uint(__uint_as_float(v) - __uint_as_float(hi)) + 0x1000u;           // rn_tf32(v - hi)
}

// bias + activation on 32 accumulator columns (= 32 output channels) of this thread's position, written transposed
// into the [32][128] staging tile (consecutive threads -> consecutive words: conflict-free)
template <int ACT>
__device__ __forceinline__ void stage_chunk_t(const uint32_t (&v)[32], const float* bias32, float* st, int r, int act_rt) {
#pragma unroll
    for (int j = 0; j < 32; ++j) st[j * kTileM + r] = activate<ACT>(__uint_as_float(v[j]) + bias32[j], act_rt);
}

// channels_last variant: the same 32 columns as one 128-byte row of this thread's position, 128B-swizzled staging tile
template <int ACT>
__device__ __forceinline__ void stage_chunk_rows(const uint32_t (&v)[32], const float* bias32, uint8_t* st, int r, int act_rt) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float4 o;
        o.x = activate<ACT>(__uint_as_float(v[4 * j + 0]) + bias32[4 * j + 0], act_rt);
        o.y = activate<ACT>(__uint_as_float(v[4 * j + 1]) + bias32[4 * j + 1], act_rt);
        o.z = activate<ACT>(__uint_as_float(v[4 * j + 2]) + bias32[4 * j + 2], act_rt);
        o.w = activate<ACT>(__uint_as_float(v[4 * j + 3]) + bias32[4 * j + 3], act_rt);
        *reinterpret_cast<float4*>(st + swz128((uint32_t)r, (uint32_t)j)) = o;
    }
}

// Work of a CTA: `full_rounds` whole items, then the last, partial round.  When that round would keep at most half of the
// CTAs busy (and the filter tile is 64 wide) its items are SPLIT along the filters, as in qnn_hamilton_tc.cu: CTA i takes
// filters [32 (i & 1), +32) of item i / 2 -- N = 32 MMAs on the same (whole) sub-filter blocks, 4 x 32 accumulator columns,
// half the output stores.  Every output element keeps its accumulation order: results are bit-identical.
struct Work {
    int item, f0, fe;
};
__device__ __forceinline__ int n_work(const P2& p) {
    return p.full_rounds + ((int)blockIdx.x < (p.split ? 2 * p.rem : p.rem) ? 1 : 0);
}
__device__ __forceinline__ Work work_of(const P2& p, int k) {
    Work w;
    if (k < p.full_rounds || !p.split) {
        w.item = (int)blockIdx.x + k * (int)gridDim.x;
        w.f0 = 0;
        w.fe = p.f_tile;
    } else {
        w.item = p.full_rounds * (int)gridDim.x + ((int)blockIdx.x >> 1);
        w.f0 = ((int)blockIdx.x & 1) * 32;
        w.fe = 32;
    }
    return w;
}

struct ItemPos {
    int ft, b, dpos, ho, w0;
};
__device__ __forceinline__ ItemPos item_pos(const P2& p, int item) {
    ItemPos ip;
    const int tile = item / p.n_ftiles;
    ip.ft = item - tile * p.n_ftiles;
    const int wt = tile % p.tiles_w, t2 = tile / p.tiles_w;
    ip.ho = t2 % p.Ho;
    const int t3 = t2 / p.Ho;
    ip.dpos = t3 % p.Do;
    ip.b = t3 / p.Do;
    ip.w0 = wt * kTileM;
    return ip;
}

template <bool CL, int ACT>
__device__ __forceinline__ void epi_phase(const uint32_t (&v)[32], int act_turn, int which, int pair, int turn, int r,
                                          int n_out, int Fp, int fe, int f0, const ItemPos& ip, const P2& p,
                                          const float* bias_s, float* st, const CUtensorMap* tmy) {
    const int c_act = pair + 2 * act_turn + 4 * which;  // 32-column chunk handled in this phase on this staging tile
    if (c_act >= n_out) return;                        // uniform across the pair
    const int col = c_act * 32;
    const int ch0 = (col / fe) * p.F + ip.ft * Fp + f0 + (col % fe);  // first of 32 consecutive output channels
    const bool mine = turn == act_turn;
    if (mine) {
        if (CL)
            stage_chunk_rows<ACT>(v, bias_s + ch0, reinterpret_cast<uint8_t*>(st), r, p.act);
        else
            stage_chunk_t<ACT>(v, bias_s + ch0, st, r, p.act);
        fence_proxy_async_smem();
    }
    named_bar_sync(5 + pair, 256);
    if (mine && r == 0) {
        if (CL)
            tma_store_4d(tmy, st, ch0, ip.w0, ip.ho, ip.b);  // y[nb][Ho][Wo][4F]: box (32 channels, 128 positions, 1, 1)
        else if (p.rank3)  // y[nb][4F][Do][Ho][Wo]: box (128 positions, 1, 1, 32 channels, 1 sample)
            tma_store_5d(tmy, st, ip.w0, ip.ho, ip.dpos, ch0, ip.b);
        else
            tma_store_4d(tmy, st, ip.w0, ip.ho, ch0, ip.b);
        tma_store_commit();
        tma_store_wait_read<0>();  // the staging tile may now be overwritten by the partner group
    }
    named_bar_sync(5 + pair, 256);
}

template <bool CONJ, bool CL, int ACT, bool X3, bool RAG>
__global__ void __launch_bounds__(kThreads, 1)
k_hamilton_tc2d(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy, const P2 p,
                const float* __restrict__ wp, const float* __restrict__ bias) {
    constexpr int kSlots = X3 ? 4 : 8;
    constexpr int kSlotCols = X3 ? 64 : 32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* b_s = smem;                                             // kSlots sub-filter blocks
    uint8_t* x_s = b_s + (size_t)kSlots * p.b_blk_bytes;             // x ring (b_blk_bytes is a multiple of 1024)
    uint8_t* y_s = x_s + (size_t)p.x_stages * p.x_stage_bytes;       // 2 staging tiles
    float* bias_s = reinterpret_cast<float*>(y_s + 2 * kStagingBytes);  // 4F floats
    Bars* bars = reinterpret_cast<Bars*>(reinterpret_cast<uint8_t*>(bias_s) + (((size_t)p.F * 16 + 1023) & ~size_t(1023)));

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int Fp = p.f_tile;

    if (tid == kWarpAlloc * 32) {
        trace(p, kTrStart);
        tma_prefetch_desc(&tmx);
        tma_prefetch_desc(&tmy);
        for (int i = 0; i < kMaxXStages; ++i) {
            mbar_init(&bars->x_full[i], 1);
            mbar_init(&bars->x_empty[i], 128);
        }
        for (int i = 0; i < kSlots; ++i) {
            mbar_init(&bars->a_full[i], 128);
            mbar_init(&bars->b_full[i], 1);
            mbar_init(&bars->a_empty[i], 4);
        }
        mbar_init(&bars->acc_full, 4);
        mbar_init(&bars->acc_empty, kEpiThreads);
        fence_mbar_init();
    }
    if (warp == kWarpAlloc) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    // Programmatic dependent launch: the prologue above may overlap the tail of the previous kernel on the stream (the
    // sub-filter pre-pass); nothing below touches global memory before that kernel has completed and flushed.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t t_acc = bars->tmem_base;
    const uint32_t t_a = t_acc + kAccCols;
    if (tid == kWarpAlloc * 32) trace(p, kTrSetup);

    uint32_t as = 0, aph = 0, accph = 0;

    if (warp >= kWarpAlloc && warp < kWarpConv0) {
        // =========================== MMA issuers (whole warp runs the loops, one lane issues) ===========================
        reg_dealloc<kRegsWg0>();
        const bool elected = elect_one();
        const int b = warp - kWarpIssuer0;  // this issuer's output component
        const uint32_t idesc_pos_f = idesc_tf32(kTileM, Fp, false, false), idesc_neg_f = idesc_tf32(kTileM, Fp, false, true);
        const uint32_t idesc_pos_h = idesc_tf32(kTileM, 32, false, false), idesc_neg_h = idesc_tf32(kTileM, 32, false, true);
        const uint64_t d0 = smem_desc_kmajor_noswz(smem_u32(b_s), (uint32_t)Fp * 16u, 128);
        const uint32_t w_lo = (uint32_t)d0, desc_hi = (uint32_t)(d0 >> 32);
        const uint32_t sub_stride = 2u * Fp;             // descriptor-lo units (16 B) between sub-filters of a block
        const uint32_t slot_stride = p.b_blk_bytes >> 4;  // ... between B slots
        const uint32_t lo_off = p.b_blk_bytes >> 5;       // 3xTF32: the lo block follows the hi block inside a slot
        constexpr uint32_t neg_table = CONJ ? kNegDense : kNegConv;
        const int slots_per_item = p.n_stages * p.KW;
        const int nw = n_work(p);
        for (int icount = 0; icount < nw; ++icount) {
            const Work wk = work_of(p, icount);
            const uint32_t idesc_pos = wk.fe == Fp ? idesc_pos_f : idesc_pos_h;
            const uint32_t idesc_neg = wk.fe == Fp ? idesc_neg_f : idesc_neg_h;
            const uint32_t d_col = t_acc + b * wk.fe;  // this issuer's accumulator: 4 x fe compact columns
            mbar_wait(&bars->acc_empty, accph ^ 1);
            tc_fence_after_sync();
            const bool tr = b == 0 && elected && icount < 12;
            if (tr) trace(p, kTrItem + 4 * icount);
            uint32_t accumulate = 0;
            for (int i = 0; i < slots_per_item; ++i) {
                mbar_wait(&bars->a_full[as], aph);
                if (tr && icount == 1 && i < 48) trace(p, kTrIssue + 2 * i);
                mbar_wait(&bars->b_full[as], aph);
                if (tr && icount == 1 && i < 48) trace(p, kTrIssue + 2 * i + 1);
                tc_fence_after_sync();
                if (elected) {
                    const uint32_t a_col = t_a + as * kSlotCols;
                    const uint32_t blk_lo = w_lo + as * slot_stride + wk.f0;  // (a filter row is 16 bytes)
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int c = a ^ b;  // sub-filter index: IDX[a][b] = a xor b
                        const uint32_t idesc = ((neg_table >> (a * 4 + b)) & 1u) ? idesc_neg : idesc_pos;
                        if (X3) {  // x_lo.w_hi + x_hi.w_lo + x_hi.w_hi
                            mma_ts(d_col, a_col + 32 + a * 8, blk_lo + c * sub_stride, desc_hi, idesc, accumulate);
                            mma_ts(d_col, a_col + a * 8, blk_lo + lo_off + c * sub_stride, desc_hi, idesc, 1);
                            mma_ts(d_col, a_col + a * 8, blk_lo + c * sub_stride, desc_hi, idesc, 1);
                        } else {
                            mma_ts(d_col, a_col + a * 8, blk_lo + c * sub_stride, desc_hi, idesc, accumulate);
                        }
                        accumulate = 1;
                    }
                    mma_commit(&bars->a_empty[as]);  // one of the four arrivals that free the A and the B slot
                }
                __syncwarp();
                if (++as == kSlots) { as = 0; aph ^= 1; }
            }
            if (elected) mma_commit(&bars->acc_full);
            if (tr) trace(p, kTrItem + 4 * icount + 1);
            __syncwarp();
            accph ^= 1;
        }
    } else if (warp >= kWarpProd0) {
        // =========================== producers: warp 28 x stages (TMA), warp 29 sub-filter blocks (bulk copies) ===========================
        reg_dealloc<kRegsWg0>();
        if (warp == kWarpProd0 && (CL ? (tid & 31) < 4 : elect_one())) {
            // channels_first: one thread, one 4-D box per stage.  channels_last: lanes 0..3 load one component each (a box
            // spanning the component axis would need a 32-byte inner row per line under a swizzle, or non-monotonic strides).
            const int comp = CL ? (tid & 31) : 0;
            uint32_t xs = 0, xph = 0;
            const int nw = n_work(p);
            for (int k = 0; k < nw; ++k) {
                const ItemPos ip = item_pos(p, work_of(p, k).item);
                const int cx = ip.w0 - p.pad_w - p.xshift, cy = ip.ho - p.pad_h, cz = ip.dpos - p.pad_d, cb = 4 * ip.b;
                for (int qc = 0; qc < p.n_qc; ++qc)
                    for (int kr = 0; kr < p.KD * p.KH; ++kr) {  // kernel planes x kernel rows: one stage each
                        const int kd = kr / p.KH, kh = kr - kd * p.KH;
                        mbar_wait(&bars->x_empty[xs], xph ^ 1);
                        if (comp == 0) mbar_arrive_expect_tx(&bars->x_full[xs], (uint32_t)((RAG ? 48 : 32) * p.wbox * 4));
                        if (CL && RAG)  // flat channel axis: box (12 channels, wbox columns, 1 row, 1 sample) from the aligned
                                        // channel at or below component * in_q + 8 qc
                            tma_load_4d(x_s + (size_t)xs * p.x_stage_bytes + (size_t)comp * p.comp_stride, &tmx, &bars->x_full[xs],
                                        (comp * p.in_q + qc * 8) & ~3, ip.w0 - p.pad_w, cy + kh * p.dh, ip.b);
                        else if (CL)  // x[nb][H][W][4][Q]: box (8 q, 1 component, wbox columns, 1 row, 1 sample) -> dense [w][8 q]
                            tma_load_5d(x_s + (size_t)xs * p.x_stage_bytes + (size_t)comp * p.comp_stride, &tmx, &bars->x_full[xs],
                                        qc * 8, comp, ip.w0 - p.pad_w, cy + kh * p.dh, ip.b);
                        else if (p.rank3)  // x[nb*4][Q][D][H][W]: box (wbox positions, 1 row, 1 plane, 8 q, 4 components)
                            tma_load_5d(x_s + (size_t)xs * p.x_stage_bytes, &tmx, &bars->x_full[xs], cx, cy + kh * p.dh,
                                        cz + kd * p.dd, qc * 8, cb);
                        else
                            tma_load_4d(x_s + (size_t)xs * p.x_stage_bytes, &tmx, &bars->x_full[xs], cx, cy + kh * p.dh, qc * 8, cb);
                        if (++xs == (uint32_t)p.x_stages) { xs = 0; xph ^= 1; }
                    }
            }
        } else if (warp == kWarpProd0 + 1 && elect_one()) {
            // block (ft, stage s, kw) of the packed sub-filters goes into the B slot paired with the A slot of (s, kw)
            const size_t blk_floats = p.b_blk_bytes >> 2;
            const int slots_per_item = p.n_stages * p.KW;
            const int nw = n_work(p);
            for (int k = 0; k < nw; ++k) {  // (a half item streams the whole blocks: the MMAs read half of their rows)
                const float* src = wp + (size_t)(work_of(p, k).item % p.n_ftiles) * slots_per_item * blk_floats;
                for (int i = 0; i < slots_per_item; ++i, src += blk_floats) {
                    mbar_wait(&bars->a_empty[as], aph ^ 1);
                    mbar_arrive_expect_tx(&bars->b_full[as], p.b_blk_bytes);
                    bulk_load(b_s + (size_t)as * p.b_blk_bytes, src, p.b_blk_bytes, &bars->b_full[as]);
                    if (++as == kSlots) { as = 0; aph ^= 1; }
                }
            }
        }
    } else if (warp >= kWarpConv0) {
        // =========================== converters: smem fp32 [channel][w] -> tf32(rn) -> TMEM A slots ===========================
        // Two groups of 128 threads take alternate x stages; a thread owns one position (TMEM lane) and walks the 32
        // channels of the stage (conflict-free 4-byte loads).  Rounding to nearest tf32 = add half an ulp, the tensor
        // core truncates.
        reg_dealloc<kRegsWg1>();
        const int cgrp = (tid - kWarpConv0 * 32) >> 7;
        const int r = (tid - kWarpConv0 * 32) & 127;
        const uint32_t lane_base = (uint32_t)(r & ~31) << 16;
        const int ch_stride = p.wbox;  // floats between channels of a stage
        int stage_i = 0;
        uint32_t xs = 0, xph = 0;
        const int my_items = n_work(p);
        const int total_stages = my_items * p.n_stages;
        for (int k = 0; k < my_items; ++k) {
            const bool tr = cgrp == 0 && r == 0 && k == 1;
            for (int s = 0; s < p.n_stages; ++s, ++stage_i) {
                if ((stage_i & 1) != cgrp) {  // the other group's stage: just advance the ring positions
                    if (++xs == (uint32_t)p.x_stages) { xs = 0; xph ^= 1; }
                    as += p.KW;
                    while (as >= kSlots) { as -= kSlots; aph ^= 1; }
                    continue;
                }
                // mbarrier parity only tells consecutive phases apart: with KW >= A slots a group's first slot wait of a stage
                // could alias a phase a whole ring older if the other group lagged, so the groups hand over explicitly -- a
                // group starts a stage only after the other one has passed every slot wait of the stage before
                // (stage_i counts over the whole kernel here: stages alternate strictly between the groups).
                if (p.handshake && stage_i > 0) named_bar_sync(12 + (cgrp ^ 1), 256);
                mbar_wait(&bars->x_full[xs], xph);
                if (tr && s < 48) trace(p, kTrConv + 4 * (s >> 1));
                const float* xb = reinterpret_cast<const float*>(x_s + (size_t)xs * p.x_stage_bytes) + r + p.xshift;
                for (int tap0 = 0; tap0 < p.KW; tap0 += kMaxTapBatch) {
                    const int nb = min(kMaxTapBatch, p.KW - tap0);
                    uint32_t as_b = as, aph_b = aph;
                    for (int tb = 0; tb < nb; ++tb) {
                        mbar_wait(&bars->a_empty[as_b], aph_b ^ 1);
                        if (++as_b == kSlots) { as_b = 0; aph_b ^= 1; }
                    }
                    tc_fence_after_sync();
                    if (p.handshake && tap0 + nb >= p.KW && stage_i + 1 < total_stages) named_bar_arrive(12 + cgrp, 256);
                    if (tr && s < 48 && tap0 == 0) trace(p, kTrConv + 4 * (s >> 1) + 1);
                    as_b = as;
                    for (int tb = 0; tb < nb; ++tb) {
                        const uint32_t dst = t_a + lane_base + as_b * kSlotCols;
                        if (CL && RAG) {
                            // stage = [component][w][12 flat channels] fp32, un-swizzled: 48 bytes per position (lanes 3
                            // 16-byte units apart: a quarter warp's loads of one unit fall on 8 different bank groups).
                            // The component's 8 channels start `shift` (0..3) channels in; channels past the component's
                            // end belong to the next one: zeroed (their weights are zero rows, but 0 x inf is NaN).
                            const uint32_t row = (uint32_t)(r + (tap0 + tb) * p.dw);
                            const uint8_t* xrow = x_s + (size_t)xs * p.x_stage_bytes + row * 48u;
                            const int qc8 = (s / (p.KD * p.KH)) * 8;
                            const int kvalid = p.in_q - qc8;  // channels of this 8-group that exist
#pragma unroll
                            for (int a = 0; a < 4; ++a) {
                                const uint4* xc = reinterpret_cast<const uint4*>(xrow + a * (uint32_t)p.comp_stride);
                                const uint4 c0 = xc[0], c1 = xc[1], c2 = xc[2];
                                uint32_t v[8];
                                switch ((a * p.in_q + qc8) & 3) {
                                    case 0: v[0] = c0.x; v[1] = c0.y; v[2] = c0.z; v[3] = c0.w; v[4] = c1.x; v[5] = c1.y; v[6] = c1.z; v[7] = c1.w; break;
                                    case 1: v[0] = c0.y; v[1] = c0.z; v[2] = c0.w; v[3] = c1.x; v[4] = c1.y; v[5] = c1.z; v[6] = c1.w; v[7] = c2.x; break;
                                    case 2: v[0] = c0.z; v[1] = c0.w; v[2] = c1.x; v[3] = c1.y; v[4] = c1.z; v[5] = c1.w; v[6] = c2.x; v[7] = c2.y; break;
                                    default: v[0] = c0.w; v[1] = c1.x; v[2] = c1.y; v[3] = c1.z; v[4] = c1.w; v[5] = c2.x; v[6] = c2.y; v[7] = c2.z; break;
                                }
#pragma unroll
                                for (int k = 0; k < 8; ++k) v[k] &= (k < kvalid ? 0xffffffffu : 0u);
                                if (X3) {
                                    uint32_t hi[8], lo[8];
#pragma unroll
                                    for (int k = 0; k < 8; ++k) split_tf32(v[k], hi[k], lo[k]);
                                    tmem_st8_nc(dst + a * 8, hi);
                                    tmem_st8_nc(dst + 32 + a * 8, lo);
                                } else {
#pragma unroll
                                    for (int k = 0; k < 8; ++k) v[k] += 0x1000u;  // round to nearest tf32
                                    tmem_st8_nc(dst + a * 8, v);
                                }
                            }
                        } else if (CL) {
                            // stage = [component][w][8 q] fp32, un-swizzled: a thread's position holds 32 bytes per component.
                            // Lanes are 32 bytes apart, so a quarter-warp's 16-byte loads of the SAME half would hit each
                            // bank twice; lanes 4..7 of every eight read the other half first -> conflict-free.
                            const uint32_t row = (uint32_t)(r + (tap0 + tb) * p.dw);
                            const uint8_t* xrow = x_s + (size_t)xs * p.x_stage_bytes + row * 32u;
                            const uint32_t comp_bytes = (uint32_t)p.comp_stride;
                            const uint32_t flip = ((uint32_t)r >> 2) & 1u;
#pragma unroll
                            for (int a2 = 0; a2 < 4; a2 += 2) {  // two components (16 columns) per store
                                uint4 q4[2][2];
#pragma unroll
                                for (int aa = 0; aa < 2; ++aa) {
                                    const uint8_t* xc = xrow + (a2 + aa) * comp_bytes;
                                    const uint4 first = *reinterpret_cast<const uint4*>(xc + (flip << 4));
                                    const uint4 second = *reinterpret_cast<const uint4*>(xc + ((flip ^ 1u) << 4));
                                    q4[aa][0] = flip ? second : first;
                                    q4[aa][1] = flip ? first : second;
                                }
                                if (X3) {
#pragma unroll
                                    for (int aa = 0; aa < 2; ++aa) {
                                        uint32_t hi[8], lo[8];
                                        split_tf32(q4[aa][0].x, hi[0], lo[0]);
                                        split_tf32(q4[aa][0].y, hi[1], lo[1]);
                                        split_tf32(q4[aa][0].z, hi[2], lo[2]);
                                        split_tf32(q4[aa][0].w, hi[3], lo[3]);
                                        split_tf32(q4[aa][1].x, hi[4], lo[4]);
                                        split_tf32(q4[aa][1].y, hi[5], lo[5]);
                                        split_tf32(q4[aa][1].z, hi[6], lo[6]);
                                        split_tf32(q4[aa][1].w, hi[7], lo[7]);
                                        tmem_st8_nc(dst + (a2 + aa) * 8, hi);
                                        tmem_st8_nc(dst + 32 + (a2 + aa) * 8, lo);
                                    }
                                } else {
                                    uint32_t u[16];
#pragma unroll
                                    for (int aa = 0; aa < 2; ++aa) {
                                        u[8 * aa + 0] = q4[aa][0].x + 0x1000u;
                                        u[8 * aa + 1] = q4[aa][0].y + 0x1000u;
                                        u[8 * aa + 2] = q4[aa][0].z + 0x1000u;
                                        u[8 * aa + 3] = q4[aa][0].w + 0x1000u;
                                        u[8 * aa + 4] = q4[aa][1].x + 0x1000u;
                                        u[8 * aa + 5] = q4[aa][1].y + 0x1000u;
                                        u[8 * aa + 6] = q4[aa][1].z + 0x1000u;
                                        u[8 * aa + 7] = q4[aa][1].w + 0x1000u;
                                    }
                                    tmem_st16_nc(dst + a2 * 8, u);
                                }
                            }
                        } else {
                            const float* xt = xb + (tap0 + tb) * p.dw;
                            if (X3) {
#pragma unroll
                                for (int g8 = 0; g8 < 4; ++g8) {  // one input component (8 channels) at a time: hi and lo
                                    uint32_t hi[8], lo[8];
#pragma unroll
                                    for (int c = 0; c < 8; ++c) split_tf32(__float_as_uint(xt[(g8 * 8 + c) * ch_stride]), hi[c], lo[c]);
                                    tmem_st8_nc(dst + g8 * 8, hi);
                                    tmem_st8_nc(dst + 32 + g8 * 8, lo);
                                }
                            } else {
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    uint32_t u[16];
#pragma unroll
                                    for (int c = 0; c < 16; ++c)
                                        u[c] = __float_as_uint(xt[(h * 16 + c) * ch_stride]) + 0x1000u;  // round to nearest tf32
                                    tmem_st16_nc(dst + h * 16, u);
                                }
                            }
                        }
                        if (++as_b == kSlots) as_b = 0;
                    }
                    if (tr && s < 48 && tap0 == 0) trace(p, kTrConv + 4 * (s >> 1) + 2);
                    tmem_wait_st();
                    tc_fence_before_sync();
                    for (int tb = 0; tb < nb; ++tb) {
                        mbar_arrive(&bars->a_full[as]);
                        if (++as == kSlots) { as = 0; aph ^= 1; }
                    }
                }
                mbar_arrive(&bars->x_empty[xs]);  // this thread is done reading the x slot
                if (tr && s < 48) trace(p, kTrConv + 4 * (s >> 1) + 3);
                if (++xs == (uint32_t)p.x_stages) { xs = 0; xph ^= 1; }
            }
        }
    } else {
        // =========================== epilogue ===========================
        reg_alloc<kRegsEpi>();
        const int e = tid;  // 0..511
        for (int i = e; i < 4 * p.F; i += kEpiThreads) bias_s[i] = p.has_bias ? __ldg(bias + i) : 0.f;
        named_bar_sync(9, kEpiThreads);
        const int grp = e >> 7, r = e & 127;
        const int pair = grp & 1, turn = grp >> 1;
        const uint32_t lane_base = (uint32_t)(r & ~31) << 16;
        float* st = reinterpret_cast<float*>(y_s + pair * kStagingBytes);
        const int nw = n_work(p);
        for (int icount = 0; icount < nw; ++icount) {
            const Work wk = work_of(p, icount);
            const int fe = wk.fe, f0 = wk.f0;
            const int n_out = (4 * fe) >> 5;  // 32-column chunks of this item: 4 or 8
            const ItemPos ip = item_pos(p, wk.item);
            mbar_wait_sleep(&bars->acc_full, accph);
            tc_fence_after_sync();
            if (e == 0 && icount < 12) trace(p, kTrItem + 4 * icount + 2);
            uint32_t v0[32], v1[32];
            tmem_ld32(t_acc + lane_base + min(grp, n_out - 1) * 32, v0);
            tmem_ld32(t_acc + lane_base + min(grp + 4, n_out - 1) * 32, v1);
            tmem_wait_ld();
            tc_fence_before_sync();
            mbar_arrive(&bars->acc_empty);  // accumulators are in registers: the next item's MMAs may start
            epi_phase<CL, ACT>(v0, 0, 0, pair, turn, r, n_out, Fp, fe, f0, ip, p, bias_s, st, &tmy);
            epi_phase<CL, ACT>(v0, 1, 0, pair, turn, r, n_out, Fp, fe, f0, ip, p, bias_s, st, &tmy);
            epi_phase<CL, ACT>(v1, 0, 1, pair, turn, r, n_out, Fp, fe, f0, ip, p, bias_s, st, &tmy);
            epi_phase<CL, ACT>(v1, 1, 1, pair, turn, r, n_out, Fp, fe, f0, ip, p, bias_s, st, &tmy);
            if (e == 0 && icount < 12) trace(p, kTrItem + 4 * icount + 3);
            accph ^= 1;
        }
        if (r == 0) tma_store_wait_read<0>();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == kWarpAlloc) tmem_dealloc(t_acc, 512);
    if (tid == kWarpAlloc * 32) trace(p, kTrEnd);
}

// Stored kernel element (tap, q, c, f) at w[tap_src * s_tap + q * s_q + c * s_c + f * s_f] (tap_src: both kernel axes flipped
// when `flip`)  ->  wp[ft][qc][tap][part][c][kg][f][q%4]  (q = qc*8 + kg*4 + q%4): the K-major, un-swizzled core-matrix image of
// each (filter tile, 8-channel chunk, tap) block, contiguous; part 0 = rn_tf32(v), part 1 (3xTF32 only) = rn_tf32(v - hi).
// The strides select the forward image or the data gradient's transposed, tap-flipped one (q and f swap roles).
__global__ void __launch_bounds__(256) k_pack_w2d(const float* __restrict__ w, float4* __restrict__ wp, int taps, int Q, int Qs,
                                                  int F, int Fp, int parts, long long s_tap, long long s_q, long long s_c,
                                                  long long s_f, int flip) {
    const int n_qc = Q >> 3, n_ft = F / Fp;
    const int total = n_ft * n_qc * taps * parts * 4 * 2 * Fp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int t = i;
        const int f = t % Fp;
        t /= Fp;
        const int kg = t & 1;
        t >>= 1;
        const int c = t & 3;
        t >>= 2;
        const int part = t % parts;
        t /= parts;
        const int tap = t % taps;
        t /= taps;
        const int qc = t % n_qc, ft = t / n_qc;
        const float* src = w + (long long)(flip ? taps - 1 - tap : tap) * s_tap + (long long)(qc * 8 + kg * 4) * s_q +
                           (long long)c * s_c + (long long)(ft * Fp + f) * s_f;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t v = (qc * 8 + kg * 4 + j < Qs) ? __float_as_uint(__ldg(src + (long long)j * s_q)) : 0u;  // rows past
            if (parts == 1) {                                                          // the layer's in_q: zero (channel padding)
                o[j] = v + 0x1000u;
            } else {
                uint32_t hi, lo;
                split_tf32(v, hi, lo);
                o[j] = part ? lo : hi;
            }
        }
        wp[i] = make_float4(__uint_as_float(o[0]), __uint_as_float(o[1]), __uint_as_float(o[2]), __uint_as_float(o[3]));
    }
}

// quaternion-channel padding for in_q % 8 != 0: x viewed as [outer][q_in][inner] -> [outer][q_out][inner], new channels
// zero.  channels_first: outer = sample x component, inner = positions; channels_last: outer = position x component,
// inner = 1.
__global__ void __launch_bounds__(256) k_pad_q(const float* __restrict__ in, float* __restrict__ out, long long outer, int q_in,
                                               int q_out, long long inner) {
    const long long total = outer * q_out * inner;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long in_i = i % inner, t = i / inner;
        const int q = (int)(t % q_out);
        const long long o = t / q_out;
        out[i] = q < q_in ? __ldg(in + (o * q_in + q) * inner + in_i) : 0.f;
    }
}

int pad_q(const float* in, float* out, long long outer, int q_in, int q_out, long long inner, cudaStream_t st) {
    const long long total = outer * q_out * inner;
    if (total == 0) return QNN_OK;
    k_pad_q<<<(unsigned)std::min<long long>((total + 255) / 256, 16LL * num_sms()), 256, 0, st>>>(in, out, outer, q_in, q_out, inner);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("channel padding launch failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    return QNN_OK;
}

// [rows][w_in] -> [rows][w_out] (w_out > w_in: zero-filled tail = padding; w_out < w_in: tail dropped)
__global__ void __launch_bounds__(256) k_copy_rows(const float* __restrict__ in, float* __restrict__ out, long long rows,
                                                   int w_in, int w_out) {
    const long long total = rows * w_out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / w_out;
        const int c = (int)(i - r * w_out);
        out[i] = c < w_in ? __ldg(in + r * w_in + c) : 0.f;
    }
}

int copy_rows(const float* in, float* out, long long rows, int w_in, int w_out, cudaStream_t st) {
    const long long total = rows * w_out;
    if (total == 0) return QNN_OK;
    k_copy_rows<<<(unsigned)std::min<long long>((total + 255) / 256, 16LL * num_sms()), 256, 0, st>>>(in, out, rows, w_in, w_out);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("row padding launch failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    return QNN_OK;
}

typedef void (*Tc2dKernel)(const CUtensorMap, const CUtensorMap, const P2, const float*, const float*);

unsigned long long* g_trace2d = nullptr;
size_t g_trace2d_bytes = 0;

template <bool CL, bool X3, bool RAG>
Tc2dKernel pick_kernel_x(int a, bool conj) {
    if (conj) return k_hamilton_tc2d<true, CL, kActLinear, X3, RAG>;  // transposed sign table: the data gradient (no activation)
    return a == kActLinear ? k_hamilton_tc2d<false, CL, kActLinear, X3, RAG>
                           : a == kActRelu ? k_hamilton_tc2d<false, CL, kActRelu, X3, RAG> : k_hamilton_tc2d<false, CL, kActGeneric, X3, RAG>;
}
Tc2dKernel pick_kernel(int act, bool channels_last, bool x3, bool conj, bool rag) {
    const int a = act == QNN_ACT_LINEAR ? kActLinear : (act == QNN_ACT_RELU ? kActRelu : kActGeneric);
    if (channels_last && rag) return x3 ? pick_kernel_x<true, true, true>(a, conj) : pick_kernel_x<true, false, true>(a, conj);
    if (channels_last) return x3 ? pick_kernel_x<true, true, false>(a, conj) : pick_kernel_x<true, false, false>(a, conj);
    return x3 ? pick_kernel_x<false, true, false>(a, conj) : pick_kernel_x<false, false, false>(a, conj);
}

}  // namespace

void tc2d_set_trace(void* device_buffer, size_t bytes) {
    g_trace2d = static_cast<unsigned long long*>(device_buffer);
    g_trace2d_bytes = bytes;
}

Tc2dPlan tc2d_plan(const Geom& g, int rank, int x3) {
    Tc2dPlan pl{};
    pl.ok = 0;
    auto no = [&](const char* why) {
        pl.why = why;
        return pl;
    };
    const bool cl = !g.channels_first;
    // channels_last rank 1 is the one-row case of rank 2 (qnn_hamilton_tc.cu, which keeps the sub-filters resident, is
    // preferred when its image fits in shared memory; this kernel streams them, so it takes what that one cannot)
    if (rank == 3 && cl) return no("channels_last rank 3 (a six-axis view: beyond the five TMA axes)");
    // the transposed sign table is instantiated without an epilogue activation: it serves the data gradient of a
    // convolution; a dense layer's own forward (conj table + activation) stays on the resident-sub-filter kernel
    if (g.conj_w && (g.dense || g.act != QNN_ACT_LINEAR)) return no("dense-layer forward (transposed sign table with an activation)");
    if (g.s[0] != 1 || g.s[1] != 1 || g.s[2] != 1) return no("stride != 1");
    // in_q % 8 != 0 (the TIMIT model's first layer has ONE quaternion input channel, interspeech_model.py:97): x goes through
    // a channel-padding pre-pass to the next multiple of 8 (zero channels; the image gets zero rows)
    // channels_last tensors with at least 8 quaternion channels are read IN PLACE instead (twelve-channel boxes from the
    // aligned channel below each component's 8-group, Tc2dPlan::rag): no pre-pass, no scratch copy of x.
    // QNN_TC2D_RAG=0 keeps the pre-pass (A/B timing)
    static const bool allow_rag = [] { const char* e = getenv("QNN_TC2D_RAG"); return !(e && atoi(e) == 0); }();
    pl.rag = (allow_rag && cl && g.in_q % 8 && g.in_q >= 8) ? 1 : 0;
    pl.pad_q = (g.in_q % 8 && !pl.rag) ? 1 : 0;
    const int Qp = (g.in_q + 7) & ~7;
    if (g.F % 32) return no("filters not a multiple of 32");
    // channels_first rows whose length is not a multiple of 4 (TMA strides must be multiples of 16 bytes; the reference's
    // TIMIT model has a free time axis, models/interspeech_model.py:81) run on row-padded scratch copies of x and y
    pl.pad_rows = (!cl && (g.in_sp[2] % 4 || g.out_sp[2] % 4)) ? 1 : 0;
    if (g.out_sp[2] < 1 || g.out_sp[1] < 1 || g.out_sp[0] < 1 || g.batch < 1) return no("empty problem");
    // channels_first: un-swizzled TMA boxes must start on a 16-byte boundary of the innermost axis (measured: an odd start
    // column is an illegal instruction, profiles/r01_tma_box_probe.log), so the box starts up to 3 columns early.
    // channels_last: the column axis is not the innermost one, any start works.
    const int xshift = cl ? 0 : (4 - (g.pad_lo[2] & 3)) & 3;
    const int wbox = cl ? kTileM + (g.k[2] - 1) * g.d[2] : (kTileM + (g.k[2] - 1) * g.d[2] + xshift + 3) & ~3;
    if (wbox > 256) return no("halo exceeds the 256-element TMA box");
    const int slots = x3 ? 4 : 8;
    const int f_tile = g.F % 64 == 0 ? 64 : 32;
    const size_t blk = (size_t)32 * f_tile * 4 * (x3 ? 2 : 1);
    // channels_last: four component blocks per stage, each starting on a 128-byte boundary (TMA destination alignment)
    const size_t comp_stride = ((size_t)wbox * (pl.rag ? 48 : 32) + 127) & ~size_t(127);
    const size_t stage = ((cl ? 4 * comp_stride : (size_t)32 * wbox * 4) + 1023) & ~size_t(1023);
    const size_t fixed = 1024 + slots * blk + 2 * kStagingBytes + (((size_t)g.F * 16 + 1023) & ~size_t(1023)) + 512;
    if (fixed + 2 * stage > kSmemLimit) return no("x stages do not fit in shared memory");
    pl.ok = 1;
    pl.f_tile = f_tile;
    pl.n_ftiles = g.F / f_tile;
    pl.wbox = wbox;
    pl.xshift = xshift;
    pl.x_stages = (int)std::min<size_t>(kMaxXStages, (kSmemLimit - fixed) / stage);
    pl.x_stage_bytes = stage;
    pl.smem_bytes = fixed + (size_t)pl.x_stages * stage;
    pl.packed_bytes = (size_t)g.k[0] * g.k[1] * g.k[2] * Qp * 4 * g.F * sizeof(float) * (x3 ? 2 : 1);
    pl.why = "";
    return pl;
}

long long tc2d_work_items(const Geom& g, const Tc2dPlan& pl) {
    const int tiles_w = (g.out_sp[2] + kTileM - 1) / kTileM;
    return (long long)g.batch * g.out_sp[0] * g.out_sp[1] * tiles_w * pl.n_ftiles;
}

size_t tc2d_packed_bytes(const Geom& g, int rank, int x3) {
    const Tc2dPlan pl = tc2d_plan(g, rank, x3);
    return pl.ok ? pl.packed_bytes : 0;
}

// `transposed`: w is the stored kernel of the layer whose DATA GRADIENT `g` describes (g.in_q = that layer's filters,
// g.F = its in_q): the image is the transposed kernel with both kernel axes flipped.
int tc2d_pack(const Geom& g, int rank, int x3, int transposed, const float* w, void* packed, cudaStream_t st) {
    const Tc2dPlan pl = tc2d_plan(g, rank, x3);
    if (!pl.ok) {
        set_error("channels_first tensor-core kernel does not take this shape: %s", pl.why);
        return QNN_E_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(packed) & 15) || !w) {
        set_error("packed kernel image must be 16-byte aligned and the kernel non-NULL");
        return QNN_E_INVALID;
    }
    const int taps = g.k[0] * g.k[1] * g.k[2], Qs = g.in_q, Q = (g.in_q + 7) & ~7, F = g.F, parts = x3 ? 2 : 1;
    long long s_tap, s_q, s_c, s_f;
    if (!transposed) {
        s_tap = (long long)Qs * 4 * F, s_q = 4LL * F, s_c = F, s_f = 1;
    } else {
        s_tap = (long long)F * 4 * Qs, s_q = 1, s_c = Qs, s_f = 4LL * Qs;
    }
    const int total = pl.n_ftiles * (Q / 8) * taps * parts * 8 * pl.f_tile;
    k_pack_w2d<<<std::min((total + 255) / 256, 4 * num_sms()), 256, 0, st>>>(w, static_cast<float4*>(packed), taps, Q, Qs, F,
                                                                          pl.f_tile, parts, s_tap, s_q, s_c, s_f,
                                                                          transposed ? 1 : 0);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("kernel packing launch failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    return QNN_OK;
}

int tc2d_forward_packed(const Geom& g, int rank, int x3, const float* x, const void* packed, const float* bias, float* y,
                        cudaStream_t st) {
    const Tc2dPlan pl = tc2d_plan(g, rank, x3);
    if (!pl.ok) {
        set_error("channels_first tensor-core kernel does not take this shape: %s", pl.why);
        return QNN_E_UNSUPPORTED;
    }
    if (pl.pad_q) {
        // in_q % 8 != 0: zero-padded quaternion channels in a scratch copy of x; the packed image already has the zero rows
        Geom gp = g;
        gp.in_q = (g.in_q + 7) & ~7;
        const long long S = (long long)g.in_sp[0] * g.in_sp[1] * g.in_sp[2];
        float* xq = nullptr;
        int rc = stream_scratch_alloc(reinterpret_cast<void**>(&xq), (size_t)g.batch * 4 * gp.in_q * S * sizeof(float), st);
        if (!rc)
            rc = g.channels_first ? pad_q(x, xq, (long long)g.batch * 4, g.in_q, gp.in_q, S, st)
                                  : pad_q(x, xq, (long long)g.batch * S * 4, g.in_q, gp.in_q, 1, st);
        if (!rc) rc = tc2d_forward_packed(gp, rank, x3, xq, packed, bias, y, st);
        if (xq) cudaFreeAsync(xq, st);
        return rc;
    }
    if (pl.pad_rows) {
        // ragged rows: x -> row-padded copy, kernel on the padded geometry (the extra input columns are zeros = the
        // convolution's own padding, the extra output columns are dropped), y <- un-padded copy
        Geom gp = g;
        gp.in_sp[2] = (g.in_sp[2] + 3) & ~3;
        gp.out_sp[2] = (g.out_sp[2] + 3) & ~3;
        const long long rows_x = (long long)g.batch * 4 * g.in_q * g.in_sp[0] * g.in_sp[1],
                        rows_y = (long long)g.batch * 4 * g.F * g.out_sp[0] * g.out_sp[1];
        float *xp = nullptr, *yp = nullptr;
        int rc = stream_scratch_alloc(reinterpret_cast<void**>(&xp), (size_t)rows_x * gp.in_sp[2] * sizeof(float), st);
        if (!rc) rc = stream_scratch_alloc(reinterpret_cast<void**>(&yp), (size_t)rows_y * gp.out_sp[2] * sizeof(float), st);
        if (!rc) rc = copy_rows(x, xp, rows_x, g.in_sp[2], gp.in_sp[2], st);
        if (!rc) rc = tc2d_forward_packed(gp, rank, x3, xp, packed, bias, yp, st);
        if (!rc) rc = copy_rows(yp, y, rows_y, gp.out_sp[2], g.out_sp[2], st);
        if (xp) cudaFreeAsync(xp, st);
        if (yp) cudaFreeAsync(yp, st);
        return rc;
    }
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(packed)) & 15) {
        set_error("tensor-core kernel needs 16-byte aligned x, packed kernel and y");
        return QNN_E_UNSUPPORTED;
    }
    const int D = g.in_sp[0], H = g.in_sp[1], W = g.in_sp[2], Do = g.out_sp[0], Ho = g.out_sp[1], Wo = g.out_sp[2], Q = g.in_q,
              F = g.F;
    P2 p{};
    p.tiles_w = (Wo + kTileM - 1) / kTileM;
    p.Ho = Ho;
    p.Do = Do;
    p.rank3 = rank == 3 ? 1 : 0;
    p.KD = g.k[0];
    p.dd = g.d[0];
    p.pad_d = g.pad_lo[0];
    p.n_ftiles = pl.n_ftiles;
    const long long ni = (long long)g.batch * Do * Ho * p.tiles_w * pl.n_ftiles;
    if (ni > 0x7fffffffLL / 64) {
        set_error("too many tiles");
        return QNN_E_UNSUPPORTED;
    }
    p.n_items = (int)ni;
    p.KH = g.k[1];
    p.KW = g.k[2];
    p.taps = g.k[0] * g.k[1] * g.k[2];
    p.dh = g.d[1];
    p.dw = g.d[2];
    p.pad_h = g.pad_lo[1];
    p.pad_w = g.pad_lo[2];
    p.n_qc = (Q + 7) / 8;
    p.rag = pl.rag;
    p.in_q = Q;
    p.n_stages = p.n_qc * p.KD * p.KH;
    p.F = F;
    p.f_tile = pl.f_tile;
    p.wbox = pl.wbox;
    p.xshift = pl.xshift;
    p.x_stages = pl.x_stages;
    p.x_stage_bytes = (int)pl.x_stage_bytes;
    p.act = g.act;
    p.has_bias = bias != nullptr;
    p.b_blk_bytes = (uint32_t)(32 * pl.f_tile * 4 * (x3 ? 2 : 1));
    p.handshake = p.KW >= (x3 ? 4 : 8) ? 1 : 0;
    p.comp_stride = (int)(((size_t)pl.wbox * (pl.rag ? 48 : 32) + 127) & ~size_t(127));

    CUtensorMap tmx, tmy;
    if (g.channels_first && rank == 3) {
        // x[nb][4][Q][D][H][W] seen as [nb*4][Q][D][H][W]: box = (wbox positions, 1 row, 1 plane, 8 quaternion channels, the
        // 4 components of one sample), no swizzle; y[nb][4F][Do][Ho][Wo]: box = (128 positions, 1, 1, 32 channels, 1 sample)
        const uint64_t dims[5] = {(uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)Q, (uint64_t)4 * g.batch};
        const uint64_t str[4] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)D * H * W * 4, (uint64_t)Q * D * H * W * 4};
        const uint32_t box[5] = {(uint32_t)pl.wbox, 1, 1, 8, 4};
        int e = make_tmap_f32(&tmx, x, 5, dims, str, box, false);
        if (e) {
            set_error("cuTensorMapEncodeTiled(x, channels_first rank 3) failed (%d)", e);
            return QNN_E_CUDA;
        }
        const uint64_t ydims[5] = {(uint64_t)Wo, (uint64_t)Ho, (uint64_t)Do, (uint64_t)4 * F, (uint64_t)g.batch};
        const uint64_t ystr[4] = {(uint64_t)Wo * 4, (uint64_t)Ho * Wo * 4, (uint64_t)Do * Ho * Wo * 4,
                                  (uint64_t)4 * F * Do * Ho * Wo * 4};
        const uint32_t ybox[5] = {(uint32_t)kTileM, 1, 1, 32, 1};
        e = make_tmap_f32(&tmy, y, 5, ydims, ystr, ybox, false);
        if (e) {
            set_error("cuTensorMapEncodeTiled(y, channels_first rank 3) failed (%d)", e);
            return QNN_E_CUDA;
        }
    } else if (g.channels_first) {
        // x[nb][4][Q][H][W] seen as [nb*4][Q][H][W]: box = (wbox positions, 1 row, 8 quaternion channels, the 4
        // components of one sample), no swizzle
        const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)Q, (uint64_t)4 * g.batch};
        const uint64_t str[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)Q * H * W * 4};
        const uint32_t box[4] = {(uint32_t)pl.wbox, 1, 8, 4};
        int e = make_tmap_f32(&tmx, x, 4, dims, str, box, false);
        if (e) {
            set_error("cuTensorMapEncodeTiled(x, channels_first) failed (%d)", e);
            return QNN_E_CUDA;
        }
        // y[nb][4F][Ho][Wo]: box = (128 positions, 1 row, 32 channels, 1 sample)
        const uint64_t ydims[4] = {(uint64_t)Wo, (uint64_t)Ho, (uint64_t)4 * F, (uint64_t)g.batch};
        const uint64_t ystr[3] = {(uint64_t)Wo * 4, (uint64_t)Ho * Wo * 4, (uint64_t)4 * F * Ho * Wo * 4};
        const uint32_t ybox[4] = {(uint32_t)kTileM, 1, 32, 1};
        e = make_tmap_f32(&tmy, y, 4, ydims, ystr, ybox, false);
        if (e) {
            set_error("cuTensorMapEncodeTiled(y, channels_first) failed (%d)", e);
            return QNN_E_CUDA;
        }
    } else {
        // x[nb][H][W][4][Q]: box = (8 quaternion channels, ONE component, wbox positions, 1 row, 1 sample), NO swizzle; four
        // such boxes (one per component, issued by four lanes) make a stage [component][w][8 q].  (One box over all four
        // components would put a 32-byte inner row on a 128-byte line each under SWIZZLE_128B -- measured,
        // profiles/r02_tma_swz_probe.log -- i.e. four times the shared memory, and 8-way bank conflicts without a swizzle.)
        const uint64_t dims[5] = {(uint64_t)Q, 4, (uint64_t)W, (uint64_t)H, (uint64_t)g.batch};
        const uint64_t str[4] = {(uint64_t)Q * 4, (uint64_t)Q * 16, (uint64_t)W * Q * 16, (uint64_t)H * W * Q * 16};
        const uint32_t box[5] = {8, 1, (uint32_t)pl.wbox, 1, 1};
        // in_q % 8 != 0 (the component blocks do not start on 16-byte boundaries): the flat channel axis, boxes of twelve
        // channels from the aligned channel at or below (component, 8-group); out-of-range channels of the last box read 0
        const uint64_t rdims[4] = {(uint64_t)Q * 4, (uint64_t)W, (uint64_t)H, (uint64_t)g.batch};
        const uint64_t rstr[3] = {(uint64_t)Q * 16, (uint64_t)W * Q * 16, (uint64_t)H * W * Q * 16};
        const uint32_t rbox[4] = {12, (uint32_t)pl.wbox, 1, 1};
        int e = pl.rag ? make_tmap_f32(&tmx, x, 4, rdims, rstr, rbox, false) : make_tmap_f32(&tmx, x, 5, dims, str, box, false);
        if (e) {
            set_error("cuTensorMapEncodeTiled(x, channels_last rank 2) failed (%d)", e);
            return QNN_E_CUDA;
        }
        // y[nb][Ho][Wo][4F]: box = (32 channels, 128 positions, 1 row, 1 sample), 128B swizzle
        const uint64_t ydims[4] = {(uint64_t)4 * F, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)g.batch};
        const uint64_t ystr[3] = {(uint64_t)F * 16, (uint64_t)Wo * F * 16, (uint64_t)Ho * Wo * F * 16};
        const uint32_t ybox[4] = {32, (uint32_t)kTileM, 1, 1};
        e = make_tmap_f32(&tmy, y, 4, ydims, ystr, ybox, true);
        if (e) {
            set_error("cuTensorMapEncodeTiled(y, channels_last rank 2) failed (%d)", e);
            return QNN_E_CUDA;
        }
    }
    Tc2dKernel kern = pick_kernel(g.act, !g.channels_first, x3 != 0, g.conj_w != 0, pl.rag != 0);
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)kSmemLimit)) return rc;
    const WorkSplit ws = plan_work_split(p.n_items, p.f_tile, num_sms());  // (qnn_common.h)
    const int grid = ws.grid;
    p.full_rounds = ws.full_rounds;
    p.rem = ws.rem;
    p.split = ws.split;
    p.trace = (g_trace2d && g_trace2d_bytes >= (size_t)grid * kTraceSlots * 8) ? g_trace2d : nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmx, tmy, p, static_cast<const float*>(packed), bias);
    count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("channels_first tensor-core kernel launch failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    return QNN_OK;
}

// Stored (un-packed) kernel: pack into stream-ordered scratch, run, release (two launches; qnn_conv_forward_packed with a
// cached image is the one-launch path).
int tc2d_forward(const Geom& g, int rank, int x3, int transposed, const float* x, const float* w, const float* bias, float* y,
                 cudaStream_t st) {
    const Tc2dPlan pl = tc2d_plan(g, rank, x3);
    if (!pl.ok) {
        set_error("channels_first tensor-core kernel does not take this shape: %s", pl.why);
        return QNN_E_UNSUPPORTED;
    }
    void* wp = nullptr;
    int rc = stream_scratch_alloc(&wp, pl.packed_bytes, st);
    if (rc) return rc;
    rc = tc2d_pack(g, rank, x3, transposed, w, wp, st);
    if (!rc) rc = tc2d_forward_packed(g, rank, x3, x, wp, bias, y, st);
    cudaFreeAsync(wp, st);
    return rc;
}

}  // namespace qnn
