// Kernel gradient of the quaternion convolution / dense layer on the 5th-generation tensor cores (sm_100a only).
//
// What TF autodiff derives from the reference graph (complexnn/conv.py:294-334 -> SURVEY 3.4): the gradient of the
// expanded 4in_q x 4F weight, folded back through concat / negate / slice into the four stored sub-filters
//     dL/df_c[tap][q][f] = sum over (a, b) with a xor b = c of  sign(a, b) * sum_pos x_a[pos + tap*dil - pad][q] * dz_b[pos][f].
// The 16 signed blocks are never formed: per k-step of 8 POSITIONS, 16 tcgen05.mma accumulate straight into FOUR
// tensor-memory accumulators D_c (rows = (tap, q), columns = f), sign = negate-B bit -- the same 4-operand sharing as
// the forward kernel with the contraction running over positions instead of channels.
//
//   A operand (x^T, rows (tap, q) x 8 positions): converter thread m owns row (tap, q) and gathers its 8 shifted positions
//       of each input component from the x stage in shared memory (TMA box: 32 + halo rows x all 4*in_q channels),
//       rounds to nearest tf32 and writes a 32-column A slot in tensor memory (4 components x 8 positions).  The tap
//       shift is part of the row index, so x is loaded once for all taps.
//   B operand (dz, 8 positions x f, K-major): the 16 packer warps read dz from global memory (coalesced 16-byte loads,
//       two sub-tiles in flight), transpose 4 x 4 blocks in registers, round to tf32 and store the core-matrix image
//       [component][position/4][f][position%4]; rows past the end of a sequence are zeros.
//   D: persistent for the whole kernel -- a CTA owns one (row block of 128 (tap, q) rows, filter tile) combination and
//       walks its share of the position sub-tiles; CTAs that own different combinations walk the same sub-tiles at the
//       same time, so x / dz come from HBM once.  At the end every CTA adds its partial sums into dW with 16-byte
//       vector reductions (red.global.add.v4.f32).
// Rank 2 (channels_last): one launch per kernel ROW kh -- for a fixed kh the problem is the rank-1 one over "sequences"
// (sample, output row ho) whose input sequence is image row ho + kh*dh - pad_h (TMA zero-fills rows outside the image),
// taps = the kernel columns; dz is shared by the launches, the launch writes dW[kh].  channels_first tensors are
// transposed to channels_last scratch copies first (qnn_api.cu).
// Warps: 0-15 packers then epilogue, 16-19 MMA issuers (one per D_c), 20-27 converters (two groups on alternate
// k-steps), 28 x producer.  TMEM: [0,256) the four accumulators, [256,512) eight A slots.
#include <algorithm>
#include <mutex>
#include "qnn_common.h"
#include "qnn_ptx.cuh"
#include "qnn_tmap.h"

namespace qnn {
namespace {
using namespace ptx;

constexpr int kSub = 32;  // positions per sub-tile (4 k-steps of 8)
constexpr int kThreads = 1024;
constexpr int kPackThreads = 512;
constexpr int kMaxSlots = 8;   // TF32: eight 32-column A slots; 3xTF32: four 64-column ones (hi | lo)
constexpr int kAccCols = 256;
constexpr int kMaxBStages = 3; // TF32: three dz stages; 3xTF32: two (each holds a hi and a lo image)
constexpr int kMaxXStages = 4;
constexpr uint32_t kSmemLimit = 232448;
constexpr int kRegsWg0 = 24, kRegsWg1 = 48, kRegsEpi = 88;
static_assert(256 * kRegsWg0 + 256 * kRegsWg1 + 512 * kRegsEpi <= 1024 * 64, "register pool");
static_assert(4 % (kSub / 8) == 0, "a unit's A slots must not wrap around the ring (4 or 8 slots)");

constexpr uint32_t kNegConv = (1u << 4) | (1u << 7) | (1u << 8) | (1u << 9) | (1u << 12) | (1u << 14);
constexpr uint32_t transpose_bits(uint32_t m) {
    uint32_t t = 0;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
            if ((m >> (a * 4 + b)) & 1u) t |= 1u << (b * 4 + a);
    return t;
}
constexpr uint32_t kNegDense = transpose_bits(kNegConv);

enum { kWarpAlloc = 16, kWarpIssuer0 = 16, kWarpConv0 = 20, kWarpProd0 = 28 };

// Optional per-CTA event trace (qnn_debug_trace, tools/wgrad_trace.py): clock64() slots per CTA.
// 0 start, 1 setup, 2 end, 3 accumulators complete, 4 epilogue done; per unit u < 24:
// 8+8u: issuer b_full passed, +1 issuer unit committed, +2 packer b_empty passed, +3 packer stored, +4 converter x_full
// passed, +5 converter unit done, +6 producer x_empty passed
constexpr int kTraceSlots = 256;

struct WP {
    unsigned long long* trace;
    int n_units, units_per_seq;  // position sub-tiles: unit -> (sample, t0 = (unit % units_per_seq) * 32)
    int n_combos, n_ftiles;      // combo = mblk * n_ftiles + ft; CTA b owns combo b % n_combos
    int taps, dil, pad_lo;
    int in_q, F, f_tile, R;      // in_q as the kernel sees x (padded to 4); R = taps * in_q rows
    int in_q_out;                // the layer's real in_q: rows with q >= in_q_out are padding and are not written
    int Lo;
    int flat_x;                  // in_q % 4 != 0 (and 4 in_q <= 256): the x box spans the flat 4 in_q channel axis (4-D map)
    int Ho, h_off;               // rank 2: sequence s = (sample n, output row ho) reads input row ho + h_off (zero when outside)
    int rows, x_stages, x_stage_bytes;
    uint32_t b_stage_bytes;
};

struct __align__(8) Bars {
    uint64_t x_full[kMaxXStages], x_empty[kMaxXStages];
    uint64_t a_full[kMaxSlots], a_empty[kMaxSlots];
    uint64_t b_full[kMaxBStages], b_empty[kMaxBStages];
    uint64_t acc_full;
    uint32_t tmem_base;
};

__device__ __forceinline__ void trace(const WP& p, int slot) {
    if (p.trace && slot < kTraceSlots) p.trace[(size_t)blockIdx.x * kTraceSlots + slot] = (unsigned long long)clock64();
}

template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

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

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void split_tf32(uint32_t v, uint32_t& hi, uint32_t& lo) {
    hi = (v + 0x1000u) & 0xffffe000u;                                          // rn_tf32(v), low bits cleared
    lo = __float_as_uint(__uint_as_float(v) - __uint_as_float(hi)) + 0x1000u;  // rn_tf32(v - hi)
}
// hi / lo parts of four values (3xTF32)
__device__ __forceinline__ void split4(float a, float b, float c, float d, float4& hi, float4& lo) {
    uint32_t h[4], l[4];
    split_tf32(__float_as_uint(a), h[0], l[0]);
    split_tf32(__float_as_uint(b), h[1], l[1]);
    split_tf32(__float_as_uint(c), h[2], l[2]);
    split_tf32(__float_as_uint(d), h[3], l[3]);
    hi = make_float4(__uint_as_float(h[0]), __uint_as_float(h[1]), __uint_as_float(h[2]), __uint_as_float(h[3]));
    lo = make_float4(__uint_as_float(l[0]), __uint_as_float(l[1]), __uint_as_float(l[2]), __uint_as_float(l[3]));
}

__device__ __forceinline__ float4 rn4(float a, float b, float c, float d) {
    return make_float4(__uint_as_float(__float_as_uint(a) + 0x1000u), __uint_as_float(__float_as_uint(b) + 0x1000u),
                       __uint_as_float(__float_as_uint(c) + 0x1000u), __uint_as_float(__float_as_uint(d) + 0x1000u));
}

// FUSED: `dz` is dy and `yfwd` the layer's relu output; the packers form dz = (y > 0 ? dy : 0) on the fly (the B operand),
// the CTAs of row block 0 also write it out for the data gradient (dz_out, may be NULL) and accumulate the bias gradient
// (db, may be NULL; pre-zeroed) -- the separate dz / bias-gradient pass over y and dy disappears.
// X3: 3xTF32 -- x^T and dz are split into hi = rn_tf32(v), lo = rn_tf32(v - hi); every block is x_lo.dz_hi + x_hi.dz_lo +
// x_hi.dz_hi (three MMAs into the same accumulator).
template <bool CONJ, bool FUSED, bool X3>
__global__ void __launch_bounds__(kThreads, 1)
k_hamilton_wgrad_tc(const __grid_constant__ CUtensorMap tmx, const WP p, const float* __restrict__ dz, float* __restrict__ dw,
                    const float* __restrict__ yfwd, float* __restrict__ dz_out, float* __restrict__ db) {
    constexpr int kSlots = X3 ? 4 : 8;
    constexpr int kSlotCols = X3 ? 64 : 32;
    constexpr int kBStages = X3 ? 2 : 3;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* b_s = smem;                                         // kBStages transposed dz sub-tiles
    uint8_t* x_s = b_s + (size_t)kBStages * p.b_stage_bytes;     // x ring
    Bars* bars = reinterpret_cast<Bars*>(x_s + (size_t)p.x_stages * p.x_stage_bytes);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int Ft = p.f_tile;
    const int combo = (int)(blockIdx.x % (unsigned)p.n_combos);
    const int cgroup = (int)(blockIdx.x / (unsigned)p.n_combos);  // which share of the units
    const int n_groups = (int)(gridDim.x / (unsigned)p.n_combos);
    const int mblk = combo / p.n_ftiles, ft = combo - mblk * p.n_ftiles;
    // CTAs beyond n_groups * n_combos (grid not a multiple of n_combos) own nothing
    const int my_units = cgroup < n_groups ? (p.n_units - cgroup + n_groups - 1) / n_groups : 0;

    if (tid == kWarpAlloc * 32) {
        trace(p, 0);
        tma_prefetch_desc(&tmx);
        for (int i = 0; i < kMaxXStages; ++i) {
            mbar_init(&bars->x_full[i], 1);
            mbar_init(&bars->x_empty[i], 256);
        }
        for (int i = 0; i < kSlots; ++i) {
            mbar_init(&bars->a_full[i], 128);
            mbar_init(&bars->a_empty[i], 4);
        }
        for (int i = 0; i < kBStages; ++i) {
            mbar_init(&bars->b_full[i], kPackThreads);
            mbar_init(&bars->b_empty[i], 4);
        }
        mbar_init(&bars->acc_full, 4);
        fence_mbar_init();
    }
    if (warp == kWarpAlloc) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t t_acc = bars->tmem_base;
    const uint32_t t_a = t_acc + kAccCols;
    if (tid == kWarpAlloc * 32) trace(p, 1);

    if (warp >= kWarpProd0) {
        // =========================== x producer ===========================
        reg_dealloc<kRegsWg0>();
        if (warp == kWarpProd0 && elect_one()) {
            uint32_t xs = 0, xph = 0;
            for (int i = 0; i < my_units; ++i) {
                const int unit = cgroup + i * n_groups;
                const int seq = unit / p.units_per_seq, t0 = (unit - seq * p.units_per_seq) * kSub;
                const int n = seq / p.Ho, ho = seq - n * p.Ho;
                mbar_wait(&bars->x_empty[xs], xph ^ 1);
                if (i < 24) trace(p, 8 + 8 * i + 6);
                mbar_arrive_expect_tx(&bars->x_full[xs], (uint32_t)(p.rows * 4 * p.in_q * 4));
                if (p.flat_x)
                    tma_load_4d(x_s + (size_t)xs * p.x_stage_bytes, &tmx, &bars->x_full[xs], 0, t0 - p.pad_lo, ho + p.h_off, n);
                else
                    tma_load_5d(x_s + (size_t)xs * p.x_stage_bytes, &tmx, &bars->x_full[xs], 0, 0, t0 - p.pad_lo, ho + p.h_off, n);
                if (++xs == (uint32_t)p.x_stages) { xs = 0; xph ^= 1; }
            }
        }
    } else if (warp >= kWarpConv0) {
        // =========================== converters: x stage -> A slots (rows (tap, q), columns 4 components x 8 positions) ===========================
        reg_dealloc<kRegsWg1>();
        const int cgrp = (tid - kWarpConv0 * 32) >> 7;
        const int m = (tid - kWarpConv0 * 32) & 127;
        const uint32_t lane_base = (uint32_t)(m & ~31) << 16;
        const int row = mblk * 128 + m;  // row of dW = tap * in_q + q
        const bool valid = row < p.R;
        const int tap = valid ? row / p.in_q : 0, q = valid ? row - tap * p.in_q : 0;
        const int pitch = 4 * p.in_q;    // floats per x row
        const int base_off = tap * p.dil * pitch + q;
        // padding lanes (row >= R) read row (tap 0, q 0) like lane 0 and mask the result: unconditional loads keep the 32
        // shared-memory reads of a k-step in flight together (predicated ones compile into 32 serialised branches)
        const uint32_t vmask = valid ? 0xffffffffu : 0u;
        uint32_t xs = 0, xph = 0, as = 0, aph = 0;
        for (int i = 0; i < my_units; ++i) {
            mbar_wait(&bars->x_full[xs], xph);
            if (cgrp == 0 && m == 0 && i < 24) trace(p, 8 + 8 * i + 4);
            const float* xb = reinterpret_cast<const float*>(x_s + (size_t)xs * p.x_stage_bytes) + base_off;
            {
                // group g converts k-steps 2g and 2g+1 of the unit as ONE batch (one tcgen05.wait::st for four stores: the
                // wait, not the store issue, is what a batch costs).  A unit's four slots never wrap (8 slots, 4 per unit).
                const uint32_t s0 = as + 2 * cgrp;
                mbar_wait(&bars->a_empty[s0], aph ^ 1);
                mbar_wait(&bars->a_empty[s0 + 1], aph ^ 1);
                tc_fence_after_sync();
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const float* xk = xb + (2 * cgrp + kk) * 8 * pitch;
                    const uint32_t dst = t_a + lane_base + (s0 + kk) * kSlotCols;
                    if (X3) {
#pragma unroll
                        for (int a = 0; a < 4; ++a) {  // one component (8 positions) at a time: hi and lo halves of the slot
                            uint32_t hi[8], lo[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                split_tf32(__float_as_uint(xk[j * pitch + a * p.in_q]), hi[j], lo[j]);
                                hi[j] &= vmask;
                                lo[j] &= vmask;
                            }
                            tmem_st8_nc(dst + a * 8, hi);
                            tmem_st8_nc(dst + 32 + a * 8, lo);
                        }
                    } else {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {  // components 2h, 2h+1: 16 columns
                            uint32_t u[16];
#pragma unroll
                            for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    u[a2 * 8 + j] = (__float_as_uint(xk[j * pitch + (2 * h + a2) * p.in_q]) + 0x1000u) & vmask;
                            tmem_st16_nc(dst + h * 16, u);
                        }
                    }
                }
                tmem_wait_st();
                tc_fence_before_sync();
                mbar_arrive(&bars->a_full[s0]);
                mbar_arrive(&bars->a_full[s0 + 1]);
                as += kSub / 8;
                if (as == kSlots) { as = 0; aph ^= 1; }
            }
            mbar_arrive(&bars->x_empty[xs]);
            if (cgrp == 0 && m == 0 && i < 24) trace(p, 8 + 8 * i + 5);
            if (++xs == (uint32_t)p.x_stages) { xs = 0; xph ^= 1; }
        }
    } else if (warp >= kWarpAlloc) {
        // =========================== MMA issuers: warp c feeds D_c ===========================
        reg_dealloc<kRegsWg0>();
        const bool elected = elect_one();
        const int c = warp - kWarpIssuer0;
        const uint32_t idesc_pos = idesc_tf32(128, Ft, false, false);
        const uint32_t idesc_neg = idesc_tf32(128, Ft, false, true);
        const uint64_t d0 = smem_desc_kmajor_noswz(smem_u32(b_s), (uint32_t)Ft * 16u, 128);
        const uint32_t b_lo = (uint32_t)d0, desc_hi = (uint32_t)(d0 >> 32);
        const uint32_t comp_stride = 8u * Ft;             // 16-byte units between the dz components of a stage
        const uint32_t kstep_stride = 2u * Ft;            // ... between k-steps (2 position groups of 4)
        const uint32_t stage_stride = p.b_stage_bytes >> 4;
        const uint32_t lo_off = p.b_stage_bytes >> 5;     // 3xTF32: the lo image follows the hi image inside a stage
        constexpr uint32_t neg_table = CONJ ? kNegDense : kNegConv;
        uint32_t as = 0, aph = 0, bs = 0, bph = 0, accumulate = 0;
        for (int i = 0; i < my_units; ++i) {
            mbar_wait(&bars->b_full[bs], bph);
            if (c == 0 && elected && i < 24) trace(p, 8 + 8 * i);
            for (int ks = 0; ks < kSub / 8; ++ks) {
                mbar_wait(&bars->a_full[as], aph);
                tc_fence_after_sync();
                if (elected) {
                    const uint32_t a_col = t_a + as * kSlotCols;
                    const uint32_t k_lo = b_lo + bs * stage_stride + ks * kstep_stride;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int b = a ^ c;
                        const uint32_t idesc = ((neg_table >> (a * 4 + b)) & 1u) ? idesc_neg : idesc_pos;
                        if (X3) {  // x_lo.dz_hi + x_hi.dz_lo + x_hi.dz_hi
                            mma_ts(t_acc + c * Ft, a_col + 32 + a * 8, k_lo + b * comp_stride, desc_hi, idesc, accumulate);
                            mma_ts(t_acc + c * Ft, a_col + a * 8, k_lo + lo_off + b * comp_stride, desc_hi, idesc, 1);
                            mma_ts(t_acc + c * Ft, a_col + a * 8, k_lo + b * comp_stride, desc_hi, idesc, 1);
                        } else {
                            mma_ts(t_acc + c * Ft, a_col + a * 8, k_lo + b * comp_stride, desc_hi, idesc, accumulate);
                        }
                        accumulate = 1;
                    }
                    mma_commit(&bars->a_empty[as]);
                }
                __syncwarp();
                if (++as == kSlots) { as = 0; aph ^= 1; }
            }
            if (elected) mma_commit(&bars->b_empty[bs]);
            if (c == 0 && elected && i < 24) trace(p, 8 + 8 * i + 1);
            __syncwarp();
            if (++bs == kBStages) { bs = 0; bph ^= 1; }
        }
        if (elected) mma_commit(&bars->acc_full);
        __syncwarp();
    } else {
        // =========================== packers (dz -> K-major tf32 blocks), then the one epilogue ===========================
        reg_alloc<kRegsEpi>();
        const int e = tid;  // 0..511
        const int f4_per = Ft >> 2;                 // 4-filter groups per component
        const int blocks = 8 * 4 * f4_per;          // 4x4 blocks per sub-tile (<= 512)
        const int fg = e % f4_per, bcomp = (e / f4_per) & 3, pg = e / (4 * f4_per);
        const bool active = e < blocks;
        const size_t C = (size_t)4 * p.F;
        const size_t col_off = (size_t)bcomp * p.F + (size_t)ft * Ft + fg * 4;
        const float* col = dz + col_off;
        float4* dst0 = reinterpret_cast<float4*>(b_s) + ((bcomp * 8 + pg) * Ft + fg * 4);
        // element offset of the first of this thread's 4 positions of unit i (rows past the end of a sequence are masked)
        auto unit_pos = [&](int i, int& t) {
            const int unit = cgroup + i * n_groups;
            const int n = unit / p.units_per_seq;
            t = (unit - n * p.units_per_seq) * kSub + pg * 4;
            return ((size_t)n * p.Lo + t) * C;
        };
        auto load_unit = [&](int i, float4 (&v)[4]) {
            int t;
            const float* src = col + unit_pos(i, t);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = (active && i < my_units && t + j < p.Lo) ? __ldg(reinterpret_cast<const float4*>(src + (size_t)j * C))
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto load_unit_y = [&](int i, float4 (&v)[4]) {  // FUSED: the forward output at the same elements
            int t;
            const float* src = yfwd + col_off + unit_pos(i, t);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = (active && i < my_units && t + j < p.Lo) ? __ldg(reinterpret_cast<const float4*>(src + (size_t)j * C))
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        uint32_t bs = 0, bph = 0;
        int ucount = 0;
        auto store_unit = [&](const float4 (&v)[4]) {
            mbar_wait(&bars->b_empty[bs], bph ^ 1);
            if (e == 0 && ucount < 24) trace(p, 8 + 8 * ucount + 2);
            if (active && X3) {
                float4* d = dst0 + (size_t)bs * (p.b_stage_bytes >> 4);
                float4* dl = d + (p.b_stage_bytes >> 5);
                const int rot = (fg >> 1) & 3;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = (i + rot) & 3;  // (see the TF32 branch for the rotation)
                    float4 hi, lo;
                    if (k == 0) split4(v[0].x, v[1].x, v[2].x, v[3].x, hi, lo);
                    else if (k == 1) split4(v[0].y, v[1].y, v[2].y, v[3].y, hi, lo);
                    else if (k == 2) split4(v[0].z, v[1].z, v[2].z, v[3].z, hi, lo);
                    else split4(v[0].w, v[1].w, v[2].w, v[3].w, hi, lo);
                    d[k] = hi;
                    dl[k] = lo;
                }
            } else if (active) {
                float4* d = dst0 + (size_t)bs * (p.b_stage_bytes >> 4);
                const float4 o0 = rn4(v[0].x, v[1].x, v[2].x, v[3].x), o1 = rn4(v[0].y, v[1].y, v[2].y, v[3].y),
                             o2 = rn4(v[0].z, v[1].z, v[2].z, v[3].z), o3 = rn4(v[0].w, v[1].w, v[2].w, v[3].w);
                // a thread owns 4 consecutive filters = four 16-byte items 64 bytes apart from its neighbour's: writing
                // item i from every lane would hit the same banks 4 ways.  Lanes rotate their order by fg / 2 instead, so
                // that 8 lanes (one 128-byte wavefront) cover 8 distinct 16-byte slots.
                const int rot = (fg >> 1) & 3;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = (i + rot) & 3;
                    d[k] = k == 0 ? o0 : k == 1 ? o1 : k == 2 ? o2 : o3;
                }
            }
            fence_proxy_async_smem();  // generic-proxy stores are read by the tensor core
            mbar_arrive(&bars->b_full[bs]);
            if (e == 0 && ucount < 24) trace(p, 8 + 8 * ucount + 3);
            ++ucount;
            if (++bs == kBStages) { bs = 0; bph ^= 1; }
        };
        if (!FUSED) {
            // three sub-tiles of dz in flight per thread (global-load latency is ~2 sub-tiles of MMA time)
            float4 v0[4], v1[4], v2[4];
            load_unit(0, v0);
            load_unit(1, v1);
            load_unit(2, v2);
            for (int i = 0; i < my_units; i += 3) {
                store_unit(v0);
                load_unit(i + 3, v0);
                if (i + 1 < my_units) {
                    store_unit(v1);
                    load_unit(i + 4, v1);
                }
                if (i + 2 < my_units) {
                    store_unit(v2);
                    load_unit(i + 5, v2);
                }
            }
        } else {
            // two sub-tiles of (dy, y) in flight per thread (the same 32 data registers per sub-tile pair as above + 16);
            // dz = relu'(y) * dy is formed when the sub-tile is packed
            const bool writer = mblk == 0 && active;  // one row block's CTAs cover every column of dz exactly once
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            auto fuse = [&](int i, float4 (&g)[4], const float4 (&yv)[4]) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g[j].x = yv[j].x > 0.f ? g[j].x : 0.f;
                    g[j].y = yv[j].y > 0.f ? g[j].y : 0.f;
                    g[j].z = yv[j].z > 0.f ? g[j].z : 0.f;
                    g[j].w = yv[j].w > 0.f ? g[j].w : 0.f;
                    acc.x += g[j].x, acc.y += g[j].y, acc.z += g[j].z, acc.w += g[j].w;
                }
                if (writer && dz_out) {
                    int t;
                    float* o = dz_out + col_off + unit_pos(i, t);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (t + j < p.Lo) *reinterpret_cast<float4*>(o + (size_t)j * C) = g[j];
                }
            };
            float4 d0[4], y0[4], d1[4], y1[4];
            load_unit(0, d0);
            load_unit_y(0, y0);
            load_unit(1, d1);
            load_unit_y(1, y1);
            for (int i = 0; i < my_units; i += 2) {
                fuse(i, d0, y0);
                store_unit(d0);
                load_unit(i + 2, d0);
                load_unit_y(i + 2, y0);
                if (i + 1 < my_units) {
                    fuse(i + 1, d1, y1);
                    store_unit(d1);
                    load_unit(i + 3, d1);
                    load_unit_y(i + 3, y1);
                }
            }
            if (db && mblk == 0) {  // (uniform per CTA) bias gradient: column sums of dz over this CTA's positions
                mbar_wait_sleep(&bars->acc_full, 0);  // every MMA has read its dz stage: the B ring is free scratch now
                float4* red = reinterpret_cast<float4*>(b_s);
                red[e] = active ? acc : make_float4(0.f, 0.f, 0.f, 0.f);
                asm volatile("bar.sync 1, %0;" ::"n"(kPackThreads) : "memory");
                if (e < 4 * f4_per) {  // e = bcomp * f4_per + fg: sum over the 8 position groups
                    float4 sum = red[e];
                    for (int g8 = 1; g8 < 8; ++g8) {
                        const int idx = g8 * 4 * f4_per + e;
                        if (idx < blocks) {
                            const float4 o = red[idx];
                            sum.x += o.x, sum.y += o.y, sum.z += o.z, sum.w += o.w;
                        }
                    }
                    float* o = db + (size_t)(e / f4_per) * p.F + (size_t)ft * Ft + (e % f4_per) * 4;
                    atomicAdd(o + 0, sum.x);
                    atomicAdd(o + 1, sum.y);
                    atomicAdd(o + 2, sum.z);
                    atomicAdd(o + 3, sum.w);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kPackThreads) : "memory");
            }
        }
        // ---- epilogue: D_c (rows (tap, q), columns f) -> dW[tap][q][c][ft*Ft + f] with vector reductions
        if (my_units > 0) {
            mbar_wait_sleep(&bars->acc_full, 0);
            tc_fence_after_sync();
            if (e == 0) trace(p, 3);
            const int m = e & 127, chunk = e >> 7;  // TMEM lane, 64-column chunk
            const uint32_t lane_base = (uint32_t)(m & ~31) << 16;
            const int row = mblk * 128 + m;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int col0 = chunk * 64 + half * 32;
                if (col0 < 4 * Ft) {  // warp-uniform
                    uint32_t v[32];
                    tmem_ld32(t_acc + lane_base + col0, v);
                    tmem_wait_ld();
                    const int tap = row / p.in_q, q = row - tap * p.in_q;
                    if (row < p.R && q < p.in_q_out) {
                        const size_t orow = (size_t)tap * p.in_q_out + q;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const int cc = col0 + j;
                            const int c = cc / Ft, f = cc - c * Ft;
                            red_add_v4(dw + (orow * 4 + c) * p.F + (size_t)ft * Ft + f, __uint_as_float(v[j]),
                                       __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        }
                    }
                }
            }
        }
    }

    if (tid == 0) trace(p, 4);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kWarpAlloc) tmem_dealloc(t_acc, 512);
    if (tid == kWarpAlloc * 32) trace(p, 2);
}

unsigned long long* g_trace_w = nullptr;
size_t g_trace_w_bytes = 0;

}  // namespace

void wgrad_set_trace(void* device_buffer, size_t bytes) {
    g_trace_w = static_cast<unsigned long long*>(device_buffer);
    g_trace_w_bytes = bytes;
}

WgradPlan wgrad_plan(const Geom& g, int rank, int x3) {
    WgradPlan pl{};
    pl.ok = 0;
    auto no = [&](const char* why) {
        pl.why = why;
        return pl;
    };
    // channels_first: the caller transposes x and dz to channels_last scratch copies (run_backward); rank 2: one launch
    // per kernel row
    if (rank > 2) return no("rank 3");
    if (g.s[1] != 1 || g.s[2] != 1) return no("stride != 1");
    if (g.in_q < 4) return no("fewer than 4 quaternion input channels");
    // in_q % 4 != 0: the component blocks of a row do not start on 16-byte boundaries, so the box cannot be (in_q, 4, rows).
    // The row as a whole (4 in_q floats = a multiple of 16 bytes) can: the x stage is then the flat row and the converter's
    // 4-byte gathers do not care about alignment -- as long as the row fits one TMA box (4 in_q <= 256).  Longer ragged rows
    // (DECODA's in_q = 250) go through the channel-padding pre-pass.
    const bool flat_x = (g.in_q % 4) != 0 && 4 * g.in_q <= 256;
    const int xq = flat_x ? g.in_q : (g.in_q + 3) & ~3;
    if (xq > 256) return no("more than 256 quaternion input channels (TMA box limit)");
    if (g.F % 16) return no("filters not a multiple of 16");
    if (g.out_sp[2] < 1 || g.out_sp[1] < 1 || g.batch < 1) return no("empty problem");
    const int taps = g.k[2];
    const int rows = kSub + (taps - 1) * g.d[2];
    if (rows > 256) return no("halo exceeds the 256-row TMA box");
    int f_tile = g.F <= 64 ? g.F : 0;  // filters per CTA (accumulators: 4 x f_tile <= 256 TMEM columns)
    for (int cand : {64, 48, 32, 16})
        if (!f_tile && g.F % cand == 0) f_tile = cand;
    const int R = taps * xq;
    const int n_mblk = (R + 127) / 128, n_ft = g.F / f_tile;
    if (n_mblk * n_ft > 148) return no("more (row block, filter tile) combinations than SMs");
    const size_t b_stage = (size_t)4 * 8 * f_tile * 16 * (x3 ? 2 : 1);
    const size_t x_stage = ((size_t)rows * 4 * xq * 4 + 1023) & ~size_t(1023);
    const size_t fixed = 1024 + (x3 ? 2 : 3) * b_stage + 512;
    if (fixed + 2 * x_stage > kSmemLimit) return no("x stages do not fit in shared memory");
    pl.ok = 1;
    pl.f_tile = f_tile;
    pl.n_ftiles = n_ft;
    pl.n_mblk = n_mblk;
    pl.rows = rows;
    pl.pad_x = xq != g.in_q;
    pl.flat_x = flat_x ? 1 : 0;
    pl.x_stages = (int)std::min<size_t>(kMaxXStages, (kSmemLimit - fixed) / x_stage);
    pl.x_stage_bytes = x_stage;
    pl.smem_bytes = fixed + (size_t)pl.x_stages * x_stage;
    pl.why = "";
    return pl;
}

// dw (stored-kernel shape [KH][KW][in_q][4][F]) is OVERWRITTEN.  x and dz are CHANNELS_LAST here whatever g.channels_first
// says (the caller hands over transposed copies): x [batch][H][W][4 in_q], dz = dy * act'(y) [batch][Ho][Wo][4F], fp32.
// Fused form (yfwd != NULL; rank 1 / dense only): `dz` is dy, yfwd the relu output of the forward; the kernel forms
// dz = relu'(y) * dy itself, writes it to dz_out (if not NULL) and adds its column sums to db (if not NULL; zeroed here).
int wgrad_tc(const Geom& g, int rank, int x3, const float* x, const float* dz, float* dw, cudaStream_t st, const float* yfwd,
             float* dz_out, float* db) {
    const WgradPlan pl = wgrad_plan(g, rank, x3);
    if (!pl.ok) {
        set_error("tensor-core kernel gradient does not take this shape: %s", pl.why);
        return QNN_E_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(dw)) & 15) {
        set_error("tensor-core kernel gradient needs 16-byte aligned x, dz and dkernel");
        return QNN_E_UNSUPPORTED;
    }
    const int H = g.in_sp[1], L = g.in_sp[2], Ho = g.out_sp[1], Lo = g.out_sp[2], KH = g.k[1], taps = g.k[2];
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)KH * taps * g.in_q * 4 * g.F * sizeof(float), st);
    if (e != cudaSuccess) {
        set_error("dkernel memset failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    const bool fused = yfwd != nullptr;
    if (fused && KH != 1) {
        set_error("the fused dz / bias-gradient form of the kernel gradient is rank 1 / dense only");
        return QNN_E_UNSUPPORTED;
    }
    if (fused && ((reinterpret_cast<uintptr_t>(yfwd) | reinterpret_cast<uintptr_t>(dz_out)) & 15)) {
        set_error("tensor-core kernel gradient needs 16-byte aligned y and dz");
        return QNN_E_UNSUPPORTED;
    }
    if (fused && db && (e = cudaMemsetAsync(db, 0, (size_t)4 * g.F * sizeof(float), st)) != cudaSuccess) {
        set_error("dbias memset failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    const int xq = pl.flat_x ? g.in_q : (g.in_q + 3) & ~3;
    float* xp = nullptr;
    if (pl.pad_x) {
        const long long rows = (long long)g.batch * H * L;
        int rc = stream_scratch_alloc(reinterpret_cast<void**>(&xp), (size_t)rows * 4 * xq * sizeof(float), st);
        if (!rc) rc = pad_x_channels(x, xp, rows, g.in_q, xq, st);
        if (rc) {
            if (xp) cudaFreeAsync(xp, st);
            return rc;
        }
        x = xp;
    }
    struct FreeOnExit {
        float* p;
        cudaStream_t st;
        ~FreeOnExit() { if (p) cudaFreeAsync(p, st); }
    } free_xp{xp, st};
    WP p{};
    p.units_per_seq = (Lo + kSub - 1) / kSub;
    const long long nu = (long long)g.batch * Ho * p.units_per_seq;
    if (nu > 0x7fffffffLL) {
        set_error("too many position sub-tiles");
        return QNN_E_UNSUPPORTED;
    }
    p.n_units = (int)nu;
    p.n_ftiles = pl.n_ftiles;
    p.n_combos = pl.n_mblk * pl.n_ftiles;
    p.taps = taps;
    p.dil = g.d[2];
    p.pad_lo = g.pad_lo[2];
    p.in_q = xq;
    p.in_q_out = g.in_q;
    p.F = g.F;
    p.f_tile = pl.f_tile;
    p.R = taps * xq;
    p.Lo = Lo;
    p.Ho = Ho;
    p.flat_x = pl.flat_x;
    p.rows = pl.rows;
    p.x_stages = pl.x_stages;
    p.x_stage_bytes = (int)pl.x_stage_bytes;
    p.b_stage_bytes = (uint32_t)(4 * 8 * pl.f_tile * 16 * (x3 ? 2 : 1));
    CUtensorMap tmx;
    if (pl.flat_x) {
        // x[nb][H][L][4 in_q]: one box = (the whole flat channel row, 32 + halo columns, one row, one sample), no swizzle
        const uint64_t dims[4] = {(uint64_t)4 * xq, (uint64_t)L, (uint64_t)H, (uint64_t)g.batch};
        const uint64_t str[3] = {(uint64_t)xq * 16, (uint64_t)L * xq * 16, (uint64_t)H * L * xq * 16};
        const uint32_t box[4] = {(uint32_t)4 * xq, (uint32_t)pl.rows, 1, 1};
        int rc = make_tmap_f32(&tmx, x, 4, dims, str, box, false);
        if (rc) {
            set_error("cuTensorMapEncodeTiled(x, wgrad, flat rows) failed (%d)", rc);
            return QNN_E_CUDA;
        }
    } else {
        // x[nb][H][L][4][in_q]: one box = (all in_q channels, the 4 components, 32 + halo columns, one row, one sample), no swizzle
        const uint64_t dims[5] = {(uint64_t)xq, 4, (uint64_t)L, (uint64_t)H, (uint64_t)g.batch};
        const uint64_t str[4] = {(uint64_t)xq * 4, (uint64_t)xq * 16, (uint64_t)L * xq * 16, (uint64_t)H * L * xq * 16};
        const uint32_t box[5] = {(uint32_t)xq, 4, (uint32_t)pl.rows, 1, 1};
        int rc = make_tmap_f32(&tmx, x, 5, dims, str, box, false);
        if (rc) {
            set_error("cuTensorMapEncodeTiled(x, wgrad) failed (%d)", rc);
            return QNN_E_CUDA;
        }
    }
    typedef void (*WKernel)(const CUtensorMap, const WP, const float*, float*, const float*, float*, float*);
    WKernel kern;
    if (x3)
        kern = fused ? (g.conj_w ? k_hamilton_wgrad_tc<true, true, true> : k_hamilton_wgrad_tc<false, true, true>)
                     : (g.conj_w ? k_hamilton_wgrad_tc<true, false, true> : k_hamilton_wgrad_tc<false, false, true>);
    else
        kern = fused ? (g.conj_w ? k_hamilton_wgrad_tc<true, true, false> : k_hamilton_wgrad_tc<false, true, false>)
                     : (g.conj_w ? k_hamilton_wgrad_tc<true, false, false> : k_hamilton_wgrad_tc<false, false, false>);
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)kSmemLimit)) return rc;
    // grid: a multiple of the number of combinations, at most one CTA per SM, no more groups than units
    int groups = std::min(num_sms() / p.n_combos, p.n_units);
    if (groups < 1) groups = 1;
    p.trace = (g_trace_w && g_trace_w_bytes >= (size_t)groups * p.n_combos * kTraceSlots * 8) ? g_trace_w : nullptr;
    for (int kh = 0; kh < KH; ++kh) {  // one launch per kernel row (rank 1: one launch)
        p.h_off = kh * g.d[1] - g.pad_lo[1];
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(groups * p.n_combos);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = pl.smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, tmx, p, dz, dw + (size_t)kh * taps * g.in_q * 4 * g.F, yfwd, dz_out, db);
        count_launch();
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_error("tensor-core kernel gradient launch failed: %s", cudaGetErrorString(e));
            return QNN_E_CUDA;
        }
    }
    return QNN_OK;
}

}  // namespace qnn
