// Fused Hamilton GEMM / 1-D convolution on the 5th-generation tensor cores (sm_100a only).
//
// Problem: rows (positions) of a channels_last tensor x[nb, L, 4*in_q], taps along L, stored un-expanded kernel
// w[taps, in_q, 4F]; y[nb, L_out, 4F] = act(bias + sum_tap sum_a sum_b  +-x_a(t + tap*dil - pad) . f_{a^b}).
// The 4in_q x 4F real weight the reference builds on every call (complexnn/conv.py:327-331, dense.py:139-143) never
// exists: the 16 signed blocks are 16 tcgen05.mma instructions that share 4 A operands (the input components) and
// 4 B operands (the sub-filters); the sign is the instruction descriptor's negate bit.
//
// Sub-filters: a pre-pass (k_pack_w1d, ~2 us, cacheable by the caller: qnn_conv_pack / qnn_*_forward_packed) writes the
// K-major, tf32-rounded core-matrix image of the stored kernel, [filter tile][hi | lo][tap][c][q/4][f][q%4]; the main
// kernel pulls the image of its filter tile into shared memory with plain bulk copies (one per lane of a warp) and keeps
// it resident.  It also is where the data gradient's transposed, tap-flipped kernel comes from (strided read).
//
// Arithmetic (template X3): TF32 = one MMA per block on operands rounded to nearest tf32; 3xTF32 = each operand is split
// into hi = rn_tf32(v) and lo = rn_tf32(v - hi) and the block is x_hi.w_hi + x_lo.w_hi + x_hi.w_lo (three MMAs, the
// dropped x_lo.w_lo term is 2^-22 relative): fp32-faithful results from the tensor cores (the reference computes in fp32,
// complexnn/conv.py:334, dense.py:149).  3xTF32 halves the A ring (4 slots of hi|lo) and doubles the resident image.
//
// One persistent CTA per SM, 896 threads, warp-specialised:
//   x tiles      TMA: raw fp32 (128 rows + halo, 32 channels) -> 128B-swizzled smem ring, issued by the converter group
//                that owns the ring slot as soon as it has consumed it (no separate producer warp).  When in_q is a
//                multiple of 8 the channel axis is walked flat (a 32-channel box may span two components, no padding);
//                otherwise per component with the out-of-range tail zero-filled by TMA.  in_q % 4 != 0: x first goes
//                through a channel-padding pre-pass (k_pad_x), TMA boxes must start on 16-byte boundaries
//   warps 20-27  converters  : two groups of four warps that take alternate x stages; smem -> registers, round-to-nearest
//                              tf32 (the tensor core would truncate), tcgen05.st into a ring of A-operand slots in tensor
//                              memory; the tap shift is a row offset in this read; all taps of a stage are converted as
//                              one batch (one tcgen05.wait::st per batch -- the wait is what a batch costs)
//   warps 16-19  MMA issuers : one warp per output component (y_r, y_i, y_j, y_k): a single thread sustains only one
//                              tcgen05.mma per ~117 cycles, four issuers reach ~37 (measured; floor 32 at N = 64);
//                              A from TMEM, B = sub-filter block resident in smem (K-major, no swizzle).  Warp 16 also
//                              initialises the barriers, fires the first x stages and the sub-filter image copies
//                              (one per lane, before anything else) and owns the TMEM allocation
//   warps 0-15   epilogue    : per tile all 16 warps pull the accumulators into registers at once (TMEM is free again
//                              after two tcgen05.ld), then +bias -> activation -> swizzled staging -> TMA store (clips
//                              ragged tiles)
// The latency-critical roles sit on the HIGHEST warp ids: the SM's issue arbiter favours high warp ids, and the 16
// epilogue warps spend most of their time polling an mbarrier (with a nanosleep back-off so they do not steal issue slots).
// TMEM columns: [0,256) four fp32 accumulators y_r|y_i|y_j|y_k (<= 64 filters per pass), [256,512) the A ring
// (TF32: eight 32-column slots; 3xTF32: four 64-column slots, hi | lo).
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
constexpr int kThreads = 896;
constexpr int kEpiThreads = 512;             // warps 0..15
constexpr int kMaxASlots = 8;
constexpr int kAccCols = 256;
constexpr int kMaxXStages = 4;
constexpr int kMaxTapBatch = 4;              // taps converted per tcgen05.wait::st (<= A slots of either mode)
constexpr int kStagingBytes = kTileM * 128;  // one [128 x 32] fp32 store tile
constexpr uint32_t kSmemLimit = 232448;      // 227 KB opt-in maximum per CTA
// register budget: 896 x 72 = 64512 at launch = 128 x kRegsWg0 + 256 x kRegsWg1 + 512 x kRegsEpi
constexpr int kRegsWg0 = 24, kRegsWg1 = 48, kRegsEpi = 96;
static_assert(128 * kRegsWg0 + 256 * kRegsWg1 + 512 * kRegsEpi <= 896 * 72, "register pool");

// bit (a*4+b) set when block (input component a -> output component b) enters negated: conv table, SURVEY 3.2.
// The dense layer uses the transposed table (bit b*4+a), SURVEY 3.3.
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
enum { kWarpAlloc = 16, kWarpIssuer0 = 16, kWarpConv0 = 20 };

// Optional per-CTA event trace (diagnostics, qnn_debug_trace): clock64() slots per CTA, see tools/tc_trace.py
constexpr int kTraceSlots = 256;  // 0..63 coarse events; 64.. detailed events of the CTA's second tile
enum { kTrStart = 0, kTrSetup = 1, kTrFirstTma = 3, kTrFirstX = 5, kTrWReady = 6,
       kTrFirstA = 7, kTrTile0 = 8 /* + 5 * tile: acc_empty passed, acc_full committed, epilogue got acc, TMEM released,
                                     stores issued */, kTrEnd = 58, kTrGlobalStart = 59, kTrGlobalEnd = 60, kTrSm = 61,
       kTrConv = 64 /* + 8*stage: x_full passed, a_empty passed (tap 0..3), wait::st done, arrived */,
       kTrIssue = 128 /* + 2*slot: a_full passed, committed */ };

struct TcParams {
    unsigned long long* trace;
    int n_tiles, tiles_per_seq;
    int taps, dil, pad_lo;
    int stride;    // output position t reads input rows t * stride + tap * dil - pad_lo
    int box_rows;  // an x stage is ceil(rows_in / box_rows) TMA boxes of box_rows rows (a box has at most 256 rows)
    int in_q, in_q_pad;
    int flat;      // 1: stages walk the flat 4*in_q channel axis (in_q % 8 == 0); 0: per component, padded to 8
    int ragged;    // in_q % 4 != 0: the component blocks of a row are not 16-byte aligned -- a stage is an un-swizzled box of
                   // 36 flat channels starting at the aligned channel at or below the chunk's first one (row pitch 144 bytes),
                   // the converter reads 4-byte words at the remaining shift (0..3 channels)
    int x_pitch;   // bytes per row of an x stage: 128 (swizzled) or 144 (ragged)
    int n_stages;  // x stages (TMA boxes) per tile
    int n_chunks;  // padded mode: 32-channel chunks per component
    int m8;        // flat mode: k-steps (8 channels) per component = in_q / 8
    int F, f_tile, n_ftiles;
    int rows_in, x_stages, x_stage_bytes;
    int act, has_bias;
    uint32_t w_bytes;      // one part (hi or lo) of one filter tile's image
    uint32_t w_img_bytes;  // what one pass keeps resident: w_bytes (TF32) or 2 * w_bytes (3xTF32: hi | lo)
    uint32_t w_chunk;      // bulk-copy granule of the image load (multiple of 16 bytes)
    int handshake;         // taps >= A slots: converter groups hand over stage by stage (see the converter role)
    int n_st;              // output staging tiles: 2 (group pairs take turns), 4 (one per epilogue group) or 8 (one per chunk)
    int full_rounds, rem;  // work of a pass: full_rounds tiles per CTA (tile = cta + k * grid), then `rem` tiles in a last round
    int split;             // the last round's tiles are split along the filters, see WorkItem
};

struct __align__(8) Barriers {
    uint64_t x_full[kMaxXStages];
    uint64_t a_full[kMaxASlots], a_empty[kMaxASlots];
    uint64_t acc_full, acc_empty, w_ready, pass_done;
    uint32_t tmem_base;
};

__device__ __forceinline__ void trace(const TcParams& p, int slot) {
    if (p.trace && slot < kTraceSlots) p.trace[(size_t)blockIdx.x * kTraceSlots + slot] = (unsigned long long)clock64();
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Register re-balancing between warpgroups (each of the four warps of a warpgroup must execute it)
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

// plain (1-D) bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Work items of a CTA in one pass: `full_rounds` whole tiles, then the last, partial round.  When that round would keep at
// most half of the CTAs busy (and the filter tile is 64 wide) its tiles are SPLIT along the filters: CTA i takes filters
// [32 (i & 1), +32) of tile i / 2 -- N = 32 MMAs into 4 x 32 accumulator columns (N = 32 issues at ~29 cycles instead of
// ~37: 0.77 of a whole tile), half the output stores (the tail of the kernel is the last tile's stores at the SM's
// ~28 B/clk write port), x read twice.  Every output element keeps its accumulation order: results are bit-identical.
struct WorkItem {
    int tile, f0, fe;  // tile, first filter of the tile's filter range, filters per component (f_tile or 32)
};
__device__ __forceinline__ int n_items(const TcParams& p) {
    return p.full_rounds + ((int)blockIdx.x < (p.split ? 2 * p.rem : p.rem) ? 1 : 0);
}
__device__ __forceinline__ WorkItem work_item(const TcParams& p, int k) {
    WorkItem w;
    if (k < p.full_rounds || !p.split) {
        w.tile = (int)blockIdx.x + k * (int)gridDim.x;
        w.f0 = 0;
        w.fe = p.f_tile;
    } else {
        w.tile = p.full_rounds * (int)gridDim.x + ((int)blockIdx.x >> 1);
        w.f0 = ((int)blockIdx.x & 1) * 32;
        w.fe = 32;
    }
    return w;
}

// Walks the x stages of a tile without integer division: for every stage the number of k-steps (8 channels each)
// and, per k-step, the input component `a` and the first quaternion channel `q0` it covers.
struct StageWalker {
    int a, j;  // flat: component and k-step index inside it;  padded: component and 32-channel chunk inside it
    __device__ __forceinline__ void reset() { a = 0; j = 0; }
    // fills ka/kq for the current stage, returns its k-step count and advances to the next stage
    __device__ __forceinline__ int next(const TcParams& p, int (&ka)[4], int (&kq)[4]) {
        if (p.flat) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                ka[ks] = a;
                kq[ks] = j << 3;
                if (++j == p.m8) { j = 0; ++a; }
            }
            return 4;
        }
        const int nks = min(32, p.in_q_pad - j * 32) >> 3;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            ka[ks] = a;
            kq[ks] = j * 32 + ks * 8;
        }
        if (++j == p.n_chunks) { j = 0; ++a; }
        return nks;
    }
};

// D[tmem] (+)= A[tmem] * B[smem]; descriptor passed as two 32-bit halves
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

// the run-time activations (tanh, sigmoid, elu, ...) as ONE out-of-line function: inlined 64 times per epilogue thread they
// made each generic-activation kernel 31 k instructions and dominated the library's compile time
__device__ __noinline__ float act_apply_call(float v, int act) { return act_apply(v, act); }

template <int ACT>
__device__ __forceinline__ float activate(float v, int act_rt) {
    if (ACT == kActLinear) return v;
    if (ACT == kActRelu) return fmaxf(v, 0.f);
    return act_apply_call(v, act_rt);
}

// Operand conversion.  TF32: "add half an ulp, let the tensor core truncate" = round to nearest, one integer add per
// element (an infinite input becomes NaN; finite inputs round exactly like cvt.rna).  3xTF32: hi = rn_tf32(v) with the
// low 13 bits cleared, lo = rn_tf32(v - hi) (the subtraction is exact in fp32).
__device__ __forceinline__ uint32_t rn_tf32(uint32_t v) { return v + 0x1000u; }
__device__ __forceinline__ void split_tf32(uint32_t v, uint32_t& hi, uint32_t& lo) {
    hi = (v + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(__uint_as_float(v) - __uint_as_float(hi)) + 0x1000u;
}

// bias + activation on 32 accumulator columns of this thread's row, written to the swizzled staging tile
template <int ACT>
__device__ __forceinline__ void stage_chunk(const uint32_t (&v)[32], const float* bias32, uint8_t* st, int r, int act_rt) {
    const float4* bs = reinterpret_cast<const float4*>(bias32);
#ifdef QNN_DIAG_NOSTAGE  // timing experiment (wrong results): the epilogue does not write its staging tiles
    if (v[0] != 0x7fc12345u) return;
#endif
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 bv = bs[j];
        float4 o;
        o.x = activate<ACT>(__uint_as_float(v[4 * j + 0]) + bv.x, act_rt);
        o.y = activate<ACT>(__uint_as_float(v[4 * j + 1]) + bv.y, act_rt);
        o.z = activate<ACT>(__uint_as_float(v[4 * j + 2]) + bv.z, act_rt);
        o.w = activate<ACT>(__uint_as_float(v[4 * j + 3]) + bv.w, act_rt);
        *reinterpret_cast<float4*>(st + swz128((uint32_t)r, (uint32_t)j)) = o;
    }
}

// One turn on a staging tile shared by the two 128-thread groups of a pair: the group whose turn it is writes its
// chunk (bias + activation), both groups meet, its thread 0 issues the TMA store and waits until the tile has been read.
template <int ACT>
__device__ __forceinline__ void epi_phase(const uint32_t (&v)[32], int act_turn, int which, int pair, int turn, int r,
                                          int n_out, int Fp, int fe, int f0, int ft, int t0, int b, const TcParams& p,
                                          const float* bias_s, uint8_t* st, const CUtensorMap* tmy) {
    const int c_act = pair + 2 * act_turn + 4 * which;  // 32-column chunk handled in this phase on this staging tile
    if (c_act >= n_out) return;                        // uniform across the pair
    const bool mine = turn == act_turn;
    // accumulator column c*32 -> component (c*32)/fe, filter f0 + (c*32)%fe of this pass' filter tile
    const int col = c_act * 32, comp = col / fe, fi = f0 + col % fe;
    if (mine) {
        stage_chunk<ACT>(v, bias_s + comp * Fp + fi, st, r, p.act);
        fence_proxy_async_smem();
    }
    named_bar_sync(5 + pair, 256);
    if (mine && r == 0) {
        tma_store_3d(tmy, st, comp * p.F + ft * Fp + fi, t0, b);
        tma_store_commit();
        tma_store_wait_read<0>();  // the staging tile may now be overwritten by the partner group
    }
    named_bar_sync(5 + pair, 256);
}

// one x stage (32 channels x rows_in rows starting at input row `row0`) = one or more row boxes into consecutive shared memory
__device__ __forceinline__ void load_x_stage(const TcParams& p, const CUtensorMap* tmx, uint8_t* dst, uint64_t* bar, int s, int row0,
                                             int b) {
    mbar_arrive_expect_tx(bar, (uint32_t)(p.rows_in * p.x_pitch));
    for (int r0 = 0; r0 < p.rows_in; r0 += p.box_rows) {  // (rows_in is a multiple of box_rows)
        if (p.flat)
            tma_load_3d(dst + (size_t)r0 * 128, tmx, bar, s * 32, row0 + r0, b);
        else if (p.ragged)  // flat channel axis, box of 36 from the aligned channel at or below component * in_q + chunk * 32
            tma_load_3d(dst + (size_t)r0 * 144, tmx, bar, ((s / p.n_chunks) * p.in_q + (s % p.n_chunks) * 32) & ~3, row0 + r0, b);
        else
            tma_load_4d(dst + (size_t)r0 * 128, tmx, bar, (s % p.n_chunks) * 32, s / p.n_chunks, row0 + r0, b);
    }
}

// the resident image of filter tile `ft`: bulk copies of w_chunk bytes, chunk `i` of `n` (the last one may be shorter)
__device__ __forceinline__ void load_w_chunk(const TcParams& p, const uint8_t* wp, uint8_t* w_s, uint64_t* bar, int ft, int i) {
    const uint32_t off = (uint32_t)i * p.w_chunk;
    const uint32_t n = min(p.w_chunk, p.w_img_bytes - off);
    bulk_load(w_s + off, wp + (size_t)ft * p.w_img_bytes + off, n, bar);
}

// RAGGED (in_q % 4 != 0) is a template parameter so that its converter path costs the common kernels nothing (as a runtime
// branch it slowed cfg 2 by 2.8 %, A/B on one box).
template <bool CONJ, int ACT, bool X3, bool RAGGED>
__global__ void __launch_bounds__(kThreads, 1)
k_hamilton_tc(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy, const TcParams p,
              const uint8_t* __restrict__ wp, const float* __restrict__ bias) {
    constexpr int kASlots = X3 ? 4 : 8;        // A ring: 256 TMEM columns either way
    constexpr int kASlotCols = X3 ? 64 : 32;   // 3xTF32: [hi (32) | lo (32)]
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (TMA 128B swizzle) by offsetting the __shared__ array itself, so that every pointer derived
    // from it stays in the shared address space (LDS/STS instead of generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* w_s = smem;                                               // resident sub-filter image of the current f-tile
    uint8_t* x_s = w_s + ((p.w_img_bytes + 1023u) & ~1023u);           // x ring
    uint8_t* y_s = x_s + (size_t)p.x_stages * p.x_stage_bytes;         // n_st staging tiles
    float* bias_s = reinterpret_cast<float*>(y_s + (size_t)p.n_st * kStagingBytes); // 4 * f_tile floats
    Barriers* bars = reinterpret_cast<Barriers*>(reinterpret_cast<uint8_t*>(bias_s) + 1024);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction
    const int Fp = p.f_tile, KQ = p.in_q_pad >> 2;
    const int n_wchunks = (int)((p.w_img_bytes + p.w_chunk - 1) / p.w_chunk);

    if (tid == kWarpAlloc * 32) {
        trace(p, kTrStart);
        if (p.trace) {
            p.trace[(size_t)blockIdx.x * kTraceSlots + kTrGlobalStart] = globaltimer_ns();
            uint32_t smid;
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            p.trace[(size_t)blockIdx.x * kTraceSlots + kTrSm] = smid;
        }
        tma_prefetch_desc(&tmx);
        tma_prefetch_desc(&tmy);
        for (int i = 0; i < kMaxXStages; ++i) {
            mbar_init(&bars->x_full[i], 1);
        }
        for (int i = 0; i < kASlots; ++i) {
            mbar_init(&bars->a_full[i], 128);
            mbar_init(&bars->a_empty[i], 4);
        }
        mbar_init(&bars->acc_full, 4);
        mbar_init(&bars->acc_empty, kEpiThreads);
        mbar_init(&bars->w_ready, 1);
        mbar_init(&bars->pass_done, kThreads);
        fence_mbar_init();
    }
    if (warp == kWarpAlloc) {
        // The first x stages and the first pass' sub-filter image are requested right here, before TMEM allocation and the
        // CTA-wide barrier, one copy per LANE of this warp (a single thread issuing all of them costs ~170 cycles
        // each): the cold HBM ramp-up -- all CTAs asking at once -- is the longest latency of the start-up.
        __syncwarp();
        asm volatile("griddepcontrol.wait;" ::: "memory");  // x / the packed kernel may be the previous kernel's output
        const int lane = tid & 31;
        const int total_stages = n_items(p) * p.n_stages;
        if (lane < p.x_stages && lane < total_stages) {
            const int j = lane / p.n_stages, s = lane - j * p.n_stages;
            const int tile = work_item(p, j).tile;
            const int b = tile / p.tiles_per_seq, t0 = (tile % p.tiles_per_seq) * kTileM;
            load_x_stage(p, &tmx, x_s + (size_t)lane * p.x_stage_bytes, &bars->x_full[lane], s, t0 * p.stride - p.pad_lo, b);
            if (lane == 0) trace(p, kTrFirstTma);
        }
        // Every CTA wants the same image at the same moment: each CTA starts at a different chunk so that the requests
        // spread over the L2 slices instead of queueing on one line set.
        if (lane == 31) mbar_arrive_expect_tx(&bars->w_ready, p.w_img_bytes);
        __syncwarp();
        for (int i = lane; i < n_wchunks; i += 32) load_w_chunk(p, wp, w_s, &bars->w_ready, 0, (i + (int)blockIdx.x) % n_wchunks);
        __syncwarp();
    }
    if (warp == kWarpAlloc) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap
    // the tail of the previous kernel on this stream; nothing below reads or writes global memory before that kernel
    // has completed and flushed.  Our own dependents may start their prologue as soon as every CTA got here.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t t_acc = bars->tmem_base;
    const uint32_t t_a = t_acc + kAccCols;
    if (tid == kWarpAlloc * 32) trace(p, kTrSetup);

    // pipeline state persists across tiles and f-tile passes
    uint32_t xs = 0, xph = 0, as = 0, aph = 0, accph = 0;

    // 896 threads x 72 registers are granted at launch.  The epilogue warpgroups hold 64 accumulator values per thread
    // and grow to 96; the producer / issuer warpgroup shrinks to 24, the two converter warpgroups to 48.
    // (setmaxnreg sits INSIDE each role branch: ptxas takes the minimum of the values that reach a join point.)

    if (warp >= kWarpAlloc && warp < kWarpConv0) {
      reg_dealloc<kRegsWg0>();
      for (int ft = 0; ft < p.n_ftiles; ++ft) {
        {
            // =========================== MMA issuers (whole warp runs the loops, one lane issues) ===========================
            const bool elected = elect_one();
            const int b = warp - kWarpIssuer0;  // this issuer's output component
            if (warp == kWarpAlloc && elected && ft > 0) {  // (the first pass' image was requested at kernel start)
                mbar_arrive_expect_tx(&bars->w_ready, p.w_img_bytes);
                int i = (int)(blockIdx.x % (unsigned)n_wchunks);
                for (int n = 0; n < n_wchunks; ++n) {
                    load_w_chunk(p, wp, w_s, &bars->w_ready, ft, i);
                    if (++i == n_wchunks) i = 0;
                }
            }
            __syncwarp();
            const uint32_t idesc_pos_f = idesc_tf32(kTileM, Fp, false, false), idesc_neg_f = idesc_tf32(kTileM, Fp, false, true);
            const uint32_t idesc_pos_h = idesc_tf32(kTileM, 32, false, false), idesc_neg_h = idesc_tf32(kTileM, 32, false, true);
            const uint64_t d0 = smem_desc_kmajor_noswz(smem_u32(w_s), (uint32_t)Fp * 16u, 128);
            const uint32_t w_lo = (uint32_t)d0, desc_hi = (uint32_t)(d0 >> 32);
            const uint32_t sub_stride = (uint32_t)KQ * Fp;  // descriptor-lo units (16 B) between sub-filters
            const uint32_t tap_stride = 4u * sub_stride;
            const uint32_t lo_off = p.w_bytes >> 4;         // 3xTF32: the lo part of the image follows the hi part
            constexpr uint32_t neg_table = CONJ ? kNegDense : kNegConv;
            mbar_wait(&bars->w_ready, ft & 1);  // this pass' image has landed (async-proxy writes, read by the tensor core)
            if (warp == kWarpIssuer0 && elected && ft == 0) trace(p, kTrWReady);
            const int items = n_items(p);
            for (int tcount = 0; tcount < items; ++tcount) {
                const WorkItem wi = work_item(p, tcount);
                const uint32_t idesc_pos = wi.fe == Fp ? idesc_pos_f : idesc_pos_h;
                const uint32_t idesc_neg = wi.fe == Fp ? idesc_neg_f : idesc_neg_h;
                const uint32_t d_col = t_acc + b * wi.fe;  // this issuer's accumulator: 4 x fe compact columns
                mbar_wait(&bars->acc_empty, accph ^ 1);
                tc_fence_after_sync();
                if (warp == kWarpIssuer0 && elected && ft == 0) trace(p, kTrTile0 + 5 * tcount);
                uint32_t accumulate = 0;
                const bool detail = warp == kWarpIssuer0 && elected && ft == 0 && tcount == 1;
                int slot_i = 0;
                StageWalker walk;
                walk.reset();
                for (int s = 0; s < p.n_stages; ++s) {
                    int ka[4], kq[4];
                    const int nks = walk.next(p, ka, kq);
                    for (int tap = 0; tap < p.taps; ++tap, ++slot_i) {
                        mbar_wait(&bars->a_full[as], aph);
                        tc_fence_after_sync();
                        if (detail && slot_i < 32) trace(p, kTrIssue + 2 * slot_i);
                        if (elected) {
                            if (warp == kWarpIssuer0 && ft == 0 && tcount == 0 && s == 0 && tap == 0) trace(p, kTrFirstA);
                            const uint32_t a_col = t_a + as * kASlotCols;
                            const uint32_t tap_lo = w_lo + tap * tap_stride + wi.f0;  // (a filter row is 16 bytes)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                if (ks >= nks) break;
                                const int a = ka[ks];
                                const int c = a ^ b;  // sub-filter index: IDX[a][b] = a xor b
                                const uint32_t k_lo = tap_lo + (uint32_t)(kq[ks] >> 2) * Fp + c * sub_stride;
                                const uint32_t idesc = ((neg_table >> (a * 4 + b)) & 1u) ? idesc_neg : idesc_pos;
                                if (X3) {
                                    // small terms first, then the leading one: x_lo.w_hi + x_hi.w_lo + x_hi.w_hi
                                    mma_ts(d_col, a_col + 32 + ks * 8, k_lo, desc_hi, idesc, accumulate);
                                    mma_ts(d_col, a_col + ks * 8, k_lo + lo_off, desc_hi, idesc, 1);
                                    mma_ts(d_col, a_col + ks * 8, k_lo, desc_hi, idesc, 1);
                                } else {
                                    mma_ts(d_col, a_col + ks * 8, k_lo, desc_hi, idesc, accumulate);
                                }
                                accumulate = 1;
                            }
                            mma_commit(&bars->a_empty[as]);  // one of the four arrivals that free the slot
                            if (detail && slot_i < 32) trace(p, kTrIssue + 2 * slot_i + 1);
                        }
                        __syncwarp();
                        if (++as == kASlots) { as = 0; aph ^= 1; }
                    }
                }
                if (elected) mma_commit(&bars->acc_full);
                if (warp == kWarpIssuer0 && elected && ft == 0) trace(p, kTrTile0 + 5 * tcount + 1);
                __syncwarp();
                accph ^= 1;
            }
        }
        // every role is done with this f-tile's sub-filters before the next pass' image overwrites them.  An mbarrier
        // (count = all threads), not bar.sync: the three roles reach it from three call sites, which compute-sanitizer's
        // synccheck reports as divergence for a hardware barrier
        mbar_arrive(&bars->pass_done);
        mbar_wait(&bars->pass_done, ft & 1);
      }
    } else if (warp >= kWarpConv0) {
      reg_dealloc<kRegsWg1>();
      for (int ft = 0; ft < p.n_ftiles; ++ft) {
        {
            // =========================== converters: smem fp32 -> tf32(rn) -> TMEM A slots ===========================
            // Two groups of 128 threads take alternate x stages, so one group's tcgen05.wait::st overlaps the other's
            // loads.  A group owns the ring slots of its parity: once its 128 threads are done with a slot, its thread 0
            // issues the TMA load of the stage that comes x_stages later into the same slot (no producer warp, no
            // "slot empty" barrier).
            const int cgrp = (tid - kWarpConv0 * 32) >> 7;
            const int r = (tid - kWarpConv0 * 32) & 127;
            const uint32_t lane_base = (uint32_t)(r & ~31) << 16;
            const int items = n_items(p);
            const int total_stages = items * p.n_stages;  // of this CTA in this pass
            // issue the x load of this pass' stage number `i` (counted over this CTA's tiles) into ring slot `slot`
            auto issue_stage = [&](int i, uint32_t slot) {
                const int j = i / p.n_stages, s = i - j * p.n_stages;
                const int tile = work_item(p, j).tile;
                const int b = tile / p.tiles_per_seq, t0 = (tile % p.tiles_per_seq) * kTileM;
                load_x_stage(p, &tmx, x_s + (size_t)slot * p.x_stage_bytes, &bars->x_full[slot], s, t0 * p.stride - p.pad_lo, b);
            };
            // prologue: fill this group's slots (x_stages is even; every pass starts at an even stage).  The first
            // pass' prologue was issued at kernel start by the barrier-initialising thread.
            if (r == 0 && ft > 0) {
                uint32_t slot = xs + cgrp;
                for (int i = cgrp; i < p.x_stages && i < total_stages; i += 2, slot += 2)
                    issue_stage(i, slot >= (uint32_t)p.x_stages ? slot - p.x_stages : slot);
            }
            int stage_i = 0;  // stage number inside this pass
            for (int it = 0; it < items; ++it) {
                const bool detail = r == 0 && ft == 0 && it == 1;
                int cch = 0, ca = 0;  // padded mode: chunk inside the component, component
                for (int s = 0; s < p.n_stages; ++s, ++stage_i) {
                    int kc = 32, shift = 0, kvalid = 32;
                    if (!p.flat) {
                        kc = min(32, p.in_q_pad - cch * 32);
                        if (RAGGED) {
                            shift = (ca * p.in_q + cch * 32) & 3;   // channels the box starts below the chunk's first one
                            kvalid = p.in_q - cch * 32;             // channels of the chunk that belong to this component
                        }
                        if (++cch == p.n_chunks) { cch = 0; ++ca; }
                    }
                    if ((stage_i & 1) != cgrp) {  // the other group's stage: just advance the ring positions
                        if (++xs == (uint32_t)p.x_stages) { xs = 0; xph ^= 1; }
                        as += p.taps;
                        while (as >= kASlots) { as -= kASlots; aph ^= 1; }
                        continue;
                    }
                    // mbarrier parity only tells consecutive phases apart.  A group's wait on the first slot of a stage is
                    // unambiguous as long as the other group's previous stage cannot lag a whole ring behind, which
                    // program order guarantees for taps < A slots.  With more taps the groups hand over explicitly: a group
                    // starts a stage only after the other one has passed every slot wait of the stage before.
                    if (p.handshake && stage_i > 0) named_bar_sync(12 + (cgrp ^ 1), 256);
                    mbar_wait(&bars->x_full[xs], xph);
                    if (detail && s < 8) trace(p, kTrConv + 8 * s);
                    if (r == 0 && stage_i == 0 && ft == 0) trace(p, kTrFirstX);
                    const uint8_t* xb = x_s + (size_t)xs * p.x_stage_bytes;
                    // (the kernel's very first stage goes tap by tap: the first MMAs start one tap's conversion after x landed)
                    const int batch = (stage_i == 0 && ft == 0) ? 1 : kMaxTapBatch;
                    for (int tap0 = 0, nb = 0; tap0 < p.taps; tap0 += nb) {
                        nb = min(batch, p.taps - tap0);
                        // a tap's stores go out as soon as ITS slot is free (the MMAs free the slots one by one, ~600 cycles
                        // apart: waiting for the whole batch first kept the first taps' stores back for no reason)
                        uint32_t as_b = as, aph_b = aph;
                        for (int tb = 0; tb < nb; ++tb) {
                            mbar_wait(&bars->a_empty[as_b], aph_b ^ 1);
                            if (detail && s < 8 && tap0 == 0) trace(p, kTrConv + 8 * s + 1 + tb);
                            tc_fence_after_sync();
                            if (p.handshake && tb == nb - 1 && tap0 + nb >= p.taps && stage_i + 1 < total_stages)
                                named_bar_arrive(12 + cgrp, 256);
                            const uint32_t row = (uint32_t)(r * p.stride + (tap0 + tb) * p.dil);
                            const uint8_t* xrow = xb + row * 128u;
                            const uint32_t sw = row & 7u;
                            const uint32_t dst = t_a + lane_base + as_b * kASlotCols;
                            if (RAGGED) {
                                // un-swizzled rows of 144 bytes; 4-byte loads at any channel shift (lanes 36 words apart:
                                // 4-way bank conflicts, the price of skipping the padding pass over x).  Channels past the
                                // component's end hold the NEXT component's data: zeroed (their weights are zero rows, but
                                // 0 x inf would still poison the sum).
                                const uint32_t* xw = reinterpret_cast<const uint32_t*>(xb + row * 144u) + shift;
                                for (int k0 = 0; k0 < kc; k0 += 8) {
                                    uint32_t v[8];
#pragma unroll
                                    for (int k = 0; k < 8; ++k)  // unconditional loads (always inside the 36-channel box), then mask
                                        v[k] = xw[k0 + k] & (k0 + k < kvalid ? 0xffffffffu : 0u);
                                    if (X3) {
                                        uint32_t hi[8], lo[8];
#pragma unroll
                                        for (int k = 0; k < 8; ++k) split_tf32(v[k], hi[k], lo[k]);
                                        tmem_st8_nc(dst + k0, hi);
                                        tmem_st8_nc(dst + 32 + k0, lo);
                                    } else {
#pragma unroll
                                        for (int k = 0; k < 8; ++k) v[k] = rn_tf32(v[k]);
                                        tmem_st8_nc(dst + k0, v);
                                    }
                                }
                            } else if (X3) {
                                for (int k0 = 0; k0 < kc; k0 += 8) {  // 8 columns at a time: hi and lo halves of the slot
                                    const uint4 v0 = *reinterpret_cast<const uint4*>(xrow + ((((k0 >> 2)) ^ sw) << 4));
                                    const uint4 v1 = *reinterpret_cast<const uint4*>(xrow + ((((k0 >> 2) + 1) ^ sw) << 4));
                                    uint32_t hi[8], lo[8];
                                    split_tf32(v0.x, hi[0], lo[0]);
                                    split_tf32(v0.y, hi[1], lo[1]);
                                    split_tf32(v0.z, hi[2], lo[2]);
                                    split_tf32(v0.w, hi[3], lo[3]);
                                    split_tf32(v1.x, hi[4], lo[4]);
                                    split_tf32(v1.y, hi[5], lo[5]);
                                    split_tf32(v1.z, hi[6], lo[6]);
                                    split_tf32(v1.w, hi[7], lo[7]);
                                    tmem_st8_nc(dst + k0, hi);
                                    tmem_st8_nc(dst + 32 + k0, lo);
                                }
                            } else if (kc == 32) {
#pragma unroll
                                for (int h = 0; h < 2; ++h) {  // 16 columns at a time: 4 loads, 16 adds, one store
                                    uint32_t u[16];
#pragma unroll
                                    for (int c4 = 0; c4 < 4; ++c4) {
#ifdef QNN_DIAG_NOLOAD  // timing experiment (wrong results): the converters do not read shared memory
                                        const uint4 v = make_uint4(row, sw, (uint32_t)(h * 4 + c4), (uint32_t)tb);
#else
                                        const uint4 v = *reinterpret_cast<const uint4*>(xrow + (((h * 4 + c4) ^ sw) << 4));
#endif
                                        u[4 * c4 + 0] = rn_tf32(v.x);
                                        u[4 * c4 + 1] = rn_tf32(v.y);
                                        u[4 * c4 + 2] = rn_tf32(v.z);
                                        u[4 * c4 + 3] = rn_tf32(v.w);
                                    }
#ifdef QNN_DIAG_NOTMEMST  // timing experiment (wrong results): the converters do not write the A slots
                                    if (u[0] == 0x7fc12345u)
#endif
                                    tmem_st16_nc(dst + h * 16, u);
                                }
                            } else {
                                for (int k0 = 0; k0 < kc; k0 += 8) {
                                    const uint4 v0 = *reinterpret_cast<const uint4*>(xrow + ((((k0 >> 2)) ^ sw) << 4));
                                    const uint4 v1 = *reinterpret_cast<const uint4*>(xrow + ((((k0 >> 2) + 1) ^ sw) << 4));
                                    const uint32_t u[8] = {rn_tf32(v0.x), rn_tf32(v0.y), rn_tf32(v0.z), rn_tf32(v0.w),
                                                           rn_tf32(v1.x), rn_tf32(v1.y), rn_tf32(v1.z), rn_tf32(v1.w)};
                                    tmem_st8_nc(dst + k0, u);
                                }
                            }
                            if (++as_b == kASlots) { as_b = 0; aph_b ^= 1; }
                        }
                        tmem_wait_st();  // one wait for the whole batch of taps
                        if (detail && s < 8 && tap0 == 0) trace(p, kTrConv + 8 * s + 5);
                        tc_fence_before_sync();
                        for (int tb = 0; tb < nb; ++tb) {
                            mbar_arrive(&bars->a_full[as]);
                            if (++as == kASlots) { as = 0; aph ^= 1; }
                        }
                        if (detail && s < 8 && tap0 == 0) trace(p, kTrConv + 8 * s + 6);
                    }
                    // every thread of the group is done reading the slot -> refill it with the stage x_stages ahead
                    named_bar_sync(10 + cgrp, 128);
                    if (r == 0 && stage_i + p.x_stages < total_stages) issue_stage(stage_i + p.x_stages, xs);
                    if (++xs == (uint32_t)p.x_stages) { xs = 0; xph ^= 1; }
                }
            }
        }
        // every role is done with this f-tile's sub-filters before the next pass' image overwrites them.  An mbarrier
        // (count = all threads), not bar.sync: the three roles reach it from three call sites, which compute-sanitizer's
        // synccheck reports as divergence for a hardware barrier
        mbar_arrive(&bars->pass_done);
        mbar_wait(&bars->pass_done, ft & 1);
      }
    } else {
      reg_alloc<kRegsEpi>();
      for (int ft = 0; ft < p.n_ftiles; ++ft) {
        {
            // =========================== epilogue ===========================
            const int e = tid;  // 0..511
            // bias of this pass, laid out like the accumulator columns: [component][f_tile]
            if (e < 4 * Fp) bias_s[e] = p.has_bias ? __ldg(bias + (e / Fp) * p.F + ft * Fp + (e % Fp)) : 0.f;  // 4 * f_tile <= 256
            named_bar_sync(9, kEpiThreads);  // bias_s visible to every epilogue thread
            // 4 groups of 128 threads (one warp per TMEM lane quadrant); group g owns 32-column chunks g and g+4.
            // Groups g and g+2 share staging tile (g & 1) and take turns on it in lock step (256-thread named barrier).
            const int grp = e >> 7, r = e & 127;
            const int pair = grp & 1, turn = grp >> 1;
            const uint32_t lane_base = (uint32_t)(r & ~31) << 16;
            uint8_t* st = y_s + pair * kStagingBytes;
            const int items = n_items(p);
            for (int tcount = 0; tcount < items; ++tcount) {
                const WorkItem wi = work_item(p, tcount);
                const int tile = wi.tile, fe = wi.fe, f0 = wi.f0;
                const int n_out = (4 * fe) >> 5;  // 32-column chunks of this item: 2, 4, 6 or 8
                const int b = tile / p.tiles_per_seq, t0 = (tile % p.tiles_per_seq) * kTileM;
                mbar_wait_sleep(&bars->acc_full, accph);  // long wait: back off, leave the issue slots to the other roles
                tc_fence_after_sync();
                if (e == 0 && ft == 0) trace(p, kTrTile0 + 5 * tcount + 2);
                // unconditional loads (clamped to a valid chunk) keep both register arrays out of local memory
                uint32_t v0[32], v1[32];
                tmem_ld32(t_acc + lane_base + min(grp, n_out - 1) * 32, v0);
                tmem_ld32(t_acc + lane_base + min(grp + 4, n_out - 1) * 32, v1);
                tmem_wait_ld();
                tc_fence_before_sync();
                mbar_arrive(&bars->acc_empty);  // accumulators are in registers: the next tile's MMAs may start
                if (e == 0 && ft == 0) trace(p, kTrTile0 + 5 * tcount + 3);
                if (p.n_st == 2 && tcount == items - 1 && ft == p.n_ftiles - 1 &&
                    (size_t)p.x_stages * p.x_stage_bytes >= 4 * (size_t)kStagingBytes) {
                    // Last tile of this CTA in the last pass: every x stage has been consumed and every MMA has read its
                    // sub-filters, so the x ring and the sub-filter region are free: each group gets two private staging
                    // tiles (no turn taking, no wait between its two chunks).
                    uint8_t* st_own[2] = {x_s + (size_t)grp * kStagingBytes, w_s + (size_t)grp * kStagingBytes};
                    const bool w_region_ok = p.w_img_bytes >= 4u * kStagingBytes;
#pragma unroll
                    for (int which = 0; which < 2; ++which) {
                        const int c = grp + 4 * which;
                        if (c >= n_out) break;
                        uint8_t* sto = (which == 1 && w_region_ok) ? st_own[1] : st_own[0];
                        if (which == 1 && !w_region_ok) {
                            if (r == 0) tma_store_wait_read<0>();
                            named_bar_sync(1 + grp, 128);
                        }
                        const int col = c * 32, comp = col / fe, fi = f0 + col % fe;
                        if (which == 0)
                            stage_chunk<ACT>(v0, bias_s + comp * Fp + fi, sto, r, p.act);
                        else
                            stage_chunk<ACT>(v1, bias_s + comp * Fp + fi, sto, r, p.act);
                        fence_proxy_async_smem();
                        named_bar_sync(1 + grp, 128);
                        if (r == 0) {
                            tma_store_3d(&tmy, sto, comp * p.F + ft * Fp + fi, t0, b);
                            tma_store_commit();
                        }
                    }
                    if (e == 0 && ft == 0) trace(p, kTrTile0 + 5 * tcount + 4);
                    accph ^= 1;
                    continue;
                }
                if (p.n_st >= 4) {
                    // Private staging: with 8 tiles every (group, chunk) owns one, with 4 every group owns one and re-uses
                    // it for its second chunk.  No turn taking between groups: the only wait is for this thread's own earlier
                    // store out of the same tile to have finished READING shared memory -- one tile period ago (8 tiles),
                    // or the chunk just issued (4 tiles).  (Round 1's two shared tiles made the epilogue 8 k cycles per
                    // tile: each chunk's store had to drain before the partner group could stage the next one.)
#pragma unroll
                    for (int which = 0; which < 2; ++which) {
                        const int c = grp + 4 * which;
                        if (c >= n_out) break;
                        uint8_t* sto = y_s + (size_t)(p.n_st == 8 ? c : grp) * kStagingBytes;
                        if (r == 0) {
                            if (which == 1 && p.n_st == 8)
                                tma_store_wait_read<1>();  // chunk 0's store of this tile may still be reading its own tile
                            else
                                tma_store_wait_read<0>();
                        }
                        named_bar_sync(1 + grp, 128);
                        const int col = c * 32, comp = col / fe, fi = f0 + col % fe;
                        if (which == 0)
                            stage_chunk<ACT>(v0, bias_s + comp * Fp + fi, sto, r, p.act);
                        else
                            stage_chunk<ACT>(v1, bias_s + comp * Fp + fi, sto, r, p.act);
                        fence_proxy_async_smem();
                        named_bar_sync(1 + grp, 128);
                        if (r == 0) {
                            tma_store_3d(&tmy, sto, comp * p.F + ft * Fp + fi, t0, b);
                            tma_store_commit();
                        }
                    }
                    if (e == 0 && ft == 0) trace(p, kTrTile0 + 5 * tcount + 4);
                    accph ^= 1;
                    continue;
                }
                // four lock-step phases on this pair's staging tile: (turn 0, chunk set 0), (1, 0), (0, 1), (1, 1)
                epi_phase<ACT>(v0, 0, 0, pair, turn, r, n_out, Fp, fe, f0, ft, t0, b, p, bias_s, st, &tmy);
                epi_phase<ACT>(v0, 1, 0, pair, turn, r, n_out, Fp, fe, f0, ft, t0, b, p, bias_s, st, &tmy);
                epi_phase<ACT>(v1, 0, 1, pair, turn, r, n_out, Fp, fe, f0, ft, t0, b, p, bias_s, st, &tmy);
                epi_phase<ACT>(v1, 1, 1, pair, turn, r, n_out, Fp, fe, f0, ft, t0, b, p, bias_s, st, &tmy);
                if (e == 0 && ft == 0) trace(p, kTrTile0 + 5 * tcount + 4);
                accph ^= 1;
            }
            if (r == 0) tma_store_wait_read<0>();  // smem may be released; completion of the writes is the kernel's end
        }
        // every role is done with this f-tile's sub-filters before the next pass' image overwrites them.  An mbarrier
        // (count = all threads), not bar.sync: the three roles reach it from three call sites, which compute-sanitizer's
        // synccheck reports as divergence for a hardware barrier
        mbar_arrive(&bars->pass_done);
        mbar_wait(&bars->pass_done, ft & 1);
      }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == kWarpAlloc) tmem_dealloc(t_acc, 512);
    if (tid == kWarpAlloc * 32) {
        trace(p, kTrEnd);
        if (p.trace) p.trace[(size_t)blockIdx.x * kTraceSlots + kTrGlobalEnd] = globaltimer_ns();
    }
}

// Stored kernel element (tap, q, c, f) at w[tap_src * s_tap + q * s_q + c * s_c + f * s_f], tap_src = flip ? taps-1-tap : tap
//   -> wp[ft][part][tap][c][q/4][f][q%4], q < in_q_pad (rows >= in_q are zero), one float4 per thread;
// part 0 = rn_tf32(v) (3xTF32: low bits cleared), part 1 (3xTF32 only) = rn_tf32(v - hi).
// The strides make this the forward image (s_q = 4F, s_c = F, s_f = 1) or the data gradient's transposed, tap-flipped
// one (the roles of q and f swapped: s_q = 1, s_f = 4F of the forward layer's kernel).
__global__ void __launch_bounds__(256) k_pack_w1d(const float* __restrict__ w, float4* __restrict__ wp, int taps, int in_q,
                                                  int KQ, int F, int Fp, int parts, long long s_tap, long long s_q,
                                                  long long s_c, long long s_f, int flip) {
    const int n_ft = F / Fp;
    const long long total = (long long)n_ft * parts * taps * 4 * KQ * Fp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int f = (int)(t % Fp);
        t /= Fp;
        const int k4 = (int)(t % KQ);
        t /= KQ;
        const int c = (int)(t & 3);
        t >>= 2;
        const int tap = (int)(t % taps);
        t /= taps;
        const int part = (int)(t % parts), ft = (int)(t / parts);
        const float* src = w + (long long)(flip ? taps - 1 - tap : tap) * s_tap + (long long)c * s_c +
                           (long long)(ft * Fp + f) * s_f;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = k4 * 4 + j;
            const uint32_t v = q < in_q ? __float_as_uint(__ldg(src + (long long)q * s_q)) : 0u;
            if (parts == 1) {
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

// x[rows][4][in_q] -> xp[rows][4][xq] (xq = in_q rounded up to 4, new channels zero): one thread per output float4
__global__ void __launch_bounds__(256) k_pad_x(const float* __restrict__ x, float4* __restrict__ xp, long long rows, int in_q,
                                               int xq) {
    const int n4 = xq >> 2;
    const long long total = rows * 4 * n4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int q0 = (int)(i % n4) * 4;
        const long long ra = i / n4;  // row * 4 + component
        const float* src = x + ra * in_q + q0;
        float4 o;
        o.x = __ldg(src);  // q0 < in_q always (xq - in_q < 4)
        o.y = q0 + 1 < in_q ? __ldg(src + 1) : 0.f;
        o.z = q0 + 2 < in_q ? __ldg(src + 2) : 0.f;
        o.w = q0 + 3 < in_q ? __ldg(src + 3) : 0.f;
        xp[i] = o;
    }
}

typedef void (*TcKernel)(const CUtensorMap, const CUtensorMap, const TcParams, const uint8_t*, const float*);

template <bool X3, bool RAGGED>
TcKernel pick_kernel_xr(bool conj, int act) {
    const int a = act == QNN_ACT_LINEAR ? kActLinear : (act == QNN_ACT_RELU ? kActRelu : kActGeneric);
    if (conj)
        return a == kActLinear ? k_hamilton_tc<true, kActLinear, X3, RAGGED>
                               : a == kActRelu ? k_hamilton_tc<true, kActRelu, X3, RAGGED> : k_hamilton_tc<true, kActGeneric, X3, RAGGED>;
    return a == kActLinear ? k_hamilton_tc<false, kActLinear, X3, RAGGED>
                           : a == kActRelu ? k_hamilton_tc<false, kActRelu, X3, RAGGED> : k_hamilton_tc<false, kActGeneric, X3, RAGGED>;
}
TcKernel pick_kernel(bool conj, int act, bool x3, bool ragged) {
    if (ragged) return x3 ? pick_kernel_xr<true, true>(conj, act) : pick_kernel_xr<false, true>(conj, act);
    return x3 ? pick_kernel_xr<true, false>(conj, act) : pick_kernel_xr<false, false>(conj, act);
}

unsigned long long* g_trace = nullptr;
size_t g_trace_bytes = 0;

cudaError_t launch_check(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return e;
}

}  // namespace

int pad_x_channels(const float* x, float* xp, long long rows, int in_q, int xq, cudaStream_t st) {
    const long long total = rows * xq;
    if (total == 0) return QNN_OK;
    k_pad_x<<<(unsigned)std::min<long long>((total + 255) / 256, 16LL * num_sms()), 256, 0, st>>>(
        x, reinterpret_cast<float4*>(xp), rows, in_q, xq);
    count_launch();
    return launch_check("channel padding") == cudaSuccess ? QNN_OK : QNN_E_CUDA;
}

void tc_set_trace(void* device_buffer, size_t bytes) {
    g_trace = static_cast<unsigned long long*>(device_buffer);
    g_trace_bytes = bytes;
}

TcPlan tc_plan(const Geom& g, int rank, int x3) {
    TcPlan pl{};
    pl.ok = 0;
    auto no = [&](const char* why) {
        pl.why = why;
        return pl;
    };
    if (g.channels_first) return no("channels_first layout");
    if (rank != 1) return no("rank > 1");
    if (g.s[2] > 4) return no("stride > 4");
    if (g.in_q < 4) return no("fewer than 4 quaternion input channels (contraction too short for the tensor cores)");
    // in_q % 4 != 0: the component blocks of x do not start on 16-byte boundaries (TMA needs that), so x goes through a
    // channel-padding pre-pass first (tc_forward); the packed kernel image is zero-padded by the pack pre-pass
    if (g.F % 16) return no("filters not a multiple of 16");
    const int taps = g.k[2];
    // input rows one tile of 128 outputs touches; more than 256 (strides > 1) are fetched as several row boxes
    int rows_in = (kTileM - 1) * g.s[2] + (taps - 1) * g.d[2] + 1, box_rows = rows_in;
    if (rows_in > 256) {
        const int n_box = (rows_in + 255) / 256;
        box_rows = ((rows_in + n_box - 1) / n_box + 7) & ~7;  // multiple of 8 rows: every box starts on a swizzle period
        rows_in = box_rows * n_box;
    }
    if (rows_in > 1024) return no("halo too large");
    if (g.out_sp[2] < 1 || g.batch < 1) return no("empty problem");
    const int in_q_pad = (g.in_q + 7) & ~7;
    const bool ragged = (g.in_q % 4) != 0;
    const size_t stage = ((size_t)rows_in * (ragged ? 144 : 128) + 1023) & ~size_t(1023);
    // Filters per pass: the whole layer when it fits (<= 64, accumulators 4 x f_tile TMEM columns); otherwise a
    // divisor that is a multiple of 32, so that every 32-column store chunk stays inside one output component.
    // 3xTF32 keeps twice the image (hi | lo) resident, so it usually settles on a smaller tile.
    int f_tile = 0, stages = 0, n_st = 2;
    size_t fixed = 0, w_bytes = 0;
    const int cand[3] = {g.F <= 64 ? g.F : 0, 64, 32};
    // Staging tiles / x stages, in order of preference.  Measured (profiles/r02_staging_sweep.json): private staging tiles
    // (4 or 8) buy nothing -- dense north-star 24.5 us with 4 + 4 vs 24.4 us with 2 + 4: the tile period is the SM's
    // store / HBM share, not the turn taking -- while dropping from 4 to 2 x stages costs 25 % (cfg 2: 31.0 -> 39.2 us).
    // So: two shared staging tiles and as many x stages as fit; QNN_TC_STAGING / QNN_TC_XSTAGES force other splits for
    // experiments.
    static const int forced_st = [] { const char* e = getenv("QNN_TC_STAGING"); return e ? atoi(e) : 0; }();
    static const int forced_xs = [] { const char* e = getenv("QNN_TC_XSTAGES"); return e ? atoi(e) : 0; }();
    const int pref[][2] = {{2, 4}, {2, 2}, {4, 4}, {4, 2}, {8, 4}, {8, 2}};
    for (int ci = 0; ci < 3 && !f_tile; ++ci) {
        const int ft = cand[ci];
        if (ft <= 0 || ft > g.F || g.F % ft || (ft != g.F && ft % 32)) continue;
        w_bytes = (size_t)taps * 4 * in_q_pad * ft * 4;
        const size_t w_pad = ((w_bytes * (x3 ? 2 : 1)) + 1023) & ~size_t(1023);
        for (const auto& pr : pref) {
            if ((forced_st && pr[0] != forced_st) || (forced_xs && pr[1] != forced_xs)) continue;
            if (!forced_st && pr[0] != 2) continue;
            const size_t fx = 1024 /*align slack*/ + w_pad + (size_t)pr[0] * kStagingBytes + 1024 /*bias*/ + 512 /*barriers*/;
            if (fx + (size_t)pr[1] * stage > kSmemLimit) continue;
            f_tile = ft;
            fixed = fx;
            n_st = pr[0];
            stages = pr[1];
            break;
        }
    }
    if (!f_tile) return no("sub-filters do not fit in shared memory for any admissible filter tile");
    pl.ok = 1;
    pl.f_tile = f_tile;
    pl.n_ftiles = g.F / f_tile;
    pl.in_q_pad = in_q_pad;
    pl.pad_x = 0;  // (round 1 padded x in a pre-pass when in_q % 4 != 0; the ragged stage mode reads x in place)
    pl.ragged = ragged ? 1 : 0;
    pl.rows_in = rows_in;
    pl.box_rows = box_rows;
    pl.x_stages = stages;
    pl.n_st = n_st;
    pl.smem_bytes = fixed + (size_t)stages * stage;
    pl.w_bytes = w_bytes;
    pl.packed_bytes = w_bytes * (x3 ? 2 : 1) * (size_t)(g.F / f_tile);
    pl.why = "";
    return pl;
}

long long tc_work_items(const Geom& g) {
    return (long long)g.batch * ((g.out_sp[2] + kTileM - 1) / kTileM);
}

size_t tc_packed_bytes(const Geom& g, int rank, int x3) {
    const TcPlan pl = tc_plan(g, rank, x3);
    return pl.ok ? pl.packed_bytes : 0;
}

// `transposed`: w is the stored kernel of the layer whose DATA GRADIENT `g` describes (g.in_q = that layer's filters,
// g.F = its in_q); the image is the transposed, tap-flipped kernel of SURVEY 3.4, read straight from the stored layout.
int tc_pack(const Geom& g, int rank, int x3, int transposed, const float* w, void* packed, cudaStream_t st) {
    const TcPlan pl = tc_plan(g, rank, x3);
    if (!pl.ok) {
        set_error("tensor-core kernel does not take this shape: %s", pl.why);
        return QNN_E_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(packed) & 15) || !w) {
        set_error("packed kernel image must be 16-byte aligned and the kernel non-NULL");
        return QNN_E_INVALID;
    }
    const int taps = g.k[2], KQ = pl.in_q_pad >> 2, parts = x3 ? 2 : 1;
    const long long total = (long long)pl.n_ftiles * parts * taps * 4 * KQ * pl.f_tile;
    long long s_tap, s_q, s_c, s_f;
    if (!transposed) {
        s_tap = (long long)g.in_q * 4 * g.F, s_q = 4LL * g.F, s_c = g.F, s_f = 1;
    } else {  // stored kernel of the forward layer: [tap][q_fwd = our f][c][f_fwd = our q]
        s_tap = (long long)g.F * 4 * g.in_q, s_q = 1, s_c = g.in_q, s_f = 4LL * g.in_q;
    }
    k_pack_w1d<<<(unsigned)std::min<long long>((total + 255) / 256, 8LL * num_sms()), 256, 0, st>>>(
        w, static_cast<float4*>(packed), taps, g.in_q, KQ, g.F, pl.f_tile, parts, s_tap, s_q, s_c, s_f, transposed ? 1 : 0);
    count_launch();
    return launch_check("kernel packing") == cudaSuccess ? QNN_OK : QNN_E_CUDA;
}

int tc_forward_packed(const Geom& g, int rank, int x3, const float* x, const void* packed, const float* bias, float* y,
                      cudaStream_t st) {
    const TcPlan pl = tc_plan(g, rank, x3);
    if (!pl.ok) {
        set_error("tensor-core kernel does not take this shape: %s", pl.why);
        return QNN_E_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(packed)) & 15) {
        set_error("tensor-core kernel needs 16-byte aligned x, packed kernel and y");
        return QNN_E_UNSUPPORTED;
    }
    const int L = g.in_sp[2], Lo = g.out_sp[2];
    const int xq = g.in_q;  // (x is read in place whatever in_q is)
    TcParams p{};
    p.tiles_per_seq = (Lo + kTileM - 1) / kTileM;
    const long long nt = (long long)g.batch * p.tiles_per_seq;
    if (nt > 0x7fffffffLL) {
        set_error("too many tiles");
        return QNN_E_UNSUPPORTED;
    }
    p.n_tiles = (int)nt;
    p.taps = g.k[2];
    p.dil = g.d[2];
    p.pad_lo = g.pad_lo[2];
    p.stride = g.s[2];
    p.box_rows = pl.box_rows;
    p.in_q = g.in_q;
    p.in_q_pad = pl.in_q_pad;
    p.flat = (xq % 8 == 0) ? 1 : 0;
    p.ragged = pl.ragged;
    p.x_pitch = pl.ragged ? 144 : 128;
    p.n_chunks = (pl.in_q_pad + 31) / 32;
    p.m8 = xq / 8;
    p.n_stages = p.flat ? xq / 8 : 4 * p.n_chunks;
    p.F = g.F;
    p.f_tile = pl.f_tile;
    p.n_ftiles = pl.n_ftiles;
    p.rows_in = pl.rows_in;
    p.x_stages = pl.x_stages;
    p.x_stage_bytes = (int)(((size_t)pl.rows_in * p.x_pitch + 1023) & ~size_t(1023));
    p.act = g.act;
    p.has_bias = bias != nullptr;
    p.w_bytes = (uint32_t)pl.w_bytes;
    p.w_img_bytes = (uint32_t)(pl.w_bytes * (x3 ? 2 : 1));
    // the image load is cut into at most 8 bulk copies (one per lane of the issuing warp; issuing a copy costs ~100 cycles
    // whatever its size -- 30 copies of 4 KB kept the warp busy for ~2.9 k cycles before the TMEM allocation), 2 KB granules
    p.w_chunk = (uint32_t)(((p.w_img_bytes + 7) / 8 + 2047) & ~2047u);
    p.handshake = p.taps >= (x3 ? 4 : 8) ? 1 : 0;
    p.n_st = pl.n_st;

    CUtensorMap tmx, tmy;
    if (p.flat) {
        const uint64_t dims[3] = {(uint64_t)xq * 4, (uint64_t)L, (uint64_t)g.batch};
        const uint64_t str[2] = {(uint64_t)xq * 16, (uint64_t)L * xq * 16};
        const uint32_t box[3] = {32, (uint32_t)pl.box_rows, 1};
        int e = make_tmap_f32(&tmx, x, 3, dims, str, box, true);
        if (e) {
            set_error("cuTensorMapEncodeTiled(x) failed (%d)", e);
            return QNN_E_CUDA;
        }
    } else if (pl.ragged) {
        // flat rows of 4 in_q floats (a multiple of 16 bytes whatever in_q is); un-swizzled boxes of 36 channels
        const uint64_t dims[3] = {(uint64_t)xq * 4, (uint64_t)L, (uint64_t)g.batch};
        const uint64_t str[2] = {(uint64_t)xq * 16, (uint64_t)L * xq * 16};
        const uint32_t box[3] = {36, (uint32_t)pl.box_rows, 1};
        int e = make_tmap_f32(&tmx, x, 3, dims, str, box, false);
        if (e) {
            set_error("cuTensorMapEncodeTiled(x, ragged) failed (%d)", e);
            return QNN_E_CUDA;
        }
    } else {
        const uint64_t dims[4] = {(uint64_t)xq, 4, (uint64_t)L, (uint64_t)g.batch};
        const uint64_t str[3] = {(uint64_t)xq * 4, (uint64_t)xq * 16, (uint64_t)L * xq * 16};
        const uint32_t box[4] = {32, 1, (uint32_t)pl.box_rows, 1};
        int e = make_tmap_f32(&tmx, x, 4, dims, str, box, true);
        if (e) {
            set_error("cuTensorMapEncodeTiled(x) failed (%d)", e);
            return QNN_E_CUDA;
        }
    }
    {
        const uint64_t dims[3] = {(uint64_t)g.F * 4, (uint64_t)Lo, (uint64_t)g.batch};
        const uint64_t str[2] = {(uint64_t)g.F * 16, (uint64_t)Lo * g.F * 16};
        const uint32_t box[3] = {32, (uint32_t)kTileM, 1};
        int e = make_tmap_f32(&tmy, y, 3, dims, str, box, true);
        if (e) {
            set_error("cuTensorMapEncodeTiled(y) failed (%d)", e);
            return QNN_E_CUDA;
        }
    }
    TcKernel kern = pick_kernel(g.conj_w != 0, g.act, x3 != 0, pl.ragged != 0);
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)kSmemLimit)) return rc;
    const WorkSplit ws = plan_work_split(p.n_tiles, p.f_tile, num_sms());  // (qnn_common.h)
    const int grid = ws.grid;
    p.full_rounds = ws.full_rounds;
    p.rem = ws.rem;
    p.split = ws.split;
    p.trace = (g_trace && g_trace_bytes >= (size_t)grid * kTraceSlots * 8) ? g_trace : nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: see griddepcontrol.wait in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmx, tmy, p, static_cast<const uint8_t*>(packed), bias);
    count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("tensor-core kernel launch failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    return QNN_OK;
}

// Stored (un-packed) kernel: pack into stream-ordered scratch, run, release -- what a caller that does not keep a packed
// image across calls gets (two launches; qnn_*_forward_packed is the one-launch path).
int tc_forward(const Geom& g, int rank, int x3, int transposed, const float* x, const float* w, const float* bias, float* y,
               cudaStream_t st) {
    const TcPlan pl = tc_plan(g, rank, x3);
    if (!pl.ok) {
        set_error("tensor-core kernel does not take this shape: %s", pl.why);
        return QNN_E_UNSUPPORTED;
    }
    void* wp = nullptr;
    int rc = stream_scratch_alloc(&wp, pl.packed_bytes, st);
    if (rc) return rc;
    rc = tc_pack(g, rank, x3, transposed, w, wp, st);
    if (!rc) rc = tc_forward_packed(g, rank, x3, x, wp, bias, y, st);
    cudaFreeAsync(wp, st);
    return rc;
}

}  // namespace qnn
