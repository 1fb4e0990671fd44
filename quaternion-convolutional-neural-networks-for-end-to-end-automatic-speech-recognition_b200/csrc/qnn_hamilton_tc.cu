// Fused Hamilton GEMM / 1-D convolution on the 5th-generation tensor cores (sm_100a only).
//
// Problem: rows (positions) of a channels_last tensor x[nb, L, 4*in_q], taps along L, stored un-expanded kernel
// w[taps, in_q, 4F]; y[nb, L_out, 4F] = act(bias + sum_tap sum_a sum_b  +-x_a(t + tap*dil - pad) . f_{a^b}).
// The 4in_q x 4F real weight the reference builds on every call (complexnn/conv.py:327-331, dense.py:139-143) never
// exists: the 16 signed blocks are 16 tcgen05.mma instructions that share 4 A operands (the input components) and
// 4 B operands (the sub-filters); the sign is the instruction descriptor's negate bit.
//
// One persistent CTA per SM, 384 threads, warp-specialised:
//   warp 0      TMA producer: raw fp32 x tiles (128 rows + halo, one component, <=32 channels) -> 128B-swizzled smem ring
//   warps 4-7   converters  : smem -> registers, round-to-nearest tf32 (the tensor core would truncate), tcgen05.st into
//                             a ring of A-operand slots in tensor memory; the tap shift is a row offset in this read
//   warp 1      MMA issuer  : per slot 4 k-steps x 4 output components, A from TMEM, B = sub-filter block resident in
//                             smem (packed + rounded once per CTA, K-major / no swizzle), accumulators in TMEM
//   warps 8-11  epilogue    : tcgen05.ld -> +bias -> activation -> swizzled staging -> TMA store (clips ragged tiles)
//   warp 2      owns the TMEM allocation
// TMEM columns: [0,256) four fp32 accumulators y_r|y_i|y_j|y_k (F <= 64 per pass), [256,512) eight 32-column A slots.
#include <algorithm>
#include <mutex>
#include "qnn_common.h"
#include "qnn_ptx.cuh"
#include "qnn_tmap.h"

namespace qnn {
namespace {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kThreads = 384;
constexpr int kASlots = 8;
constexpr int kASlotCols = 32;
constexpr int kAccCols = 256;
constexpr int kMaxXStages = 4;
constexpr int kStagingBytes = kTileM * 128;  // one [128 x 32] fp32 store tile
constexpr uint32_t kSmemLimit = 232448;      // 227 KB opt-in maximum per CTA

// bit (a*4+b) set when block (input component a -> output component b) enters negated: conv table, SURVEY 3.2
constexpr uint32_t kNegConv = (1u << 4) | (1u << 7) | (1u << 8) | (1u << 9) | (1u << 12) | (1u << 14);

struct TcParams {
    int n_tiles, tiles_per_seq;
    int taps, dil, pad_lo;
    int in_q, in_q_pad, n_chunks;
    int F, f_tile, n_ftiles;
    int rows_in, x_stages, x_stage_bytes;
    int act, conj_w, has_bias;
    uint32_t w_bytes;
};

struct __align__(8) Barriers {
    uint64_t x_full[kMaxXStages], x_empty[kMaxXStages];
    uint64_t a_full[kASlots], a_empty[kASlots];
    uint64_t acc_full, acc_empty;
    uint32_t tmem_base;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
k_hamilton_tc(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy, const TcParams p,
              const float* __restrict__ w, const float* __restrict__ bias) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w_s = smem;                                               // resident sub-filters of the current f-tile
    uint8_t* x_s = w_s + ((p.w_bytes + 1023u) & ~1023u);               // x ring
    uint8_t* y_s = x_s + (size_t)p.x_stages * p.x_stage_bytes;         // 2 staging tiles
    float* bias_s = reinterpret_cast<float*>(y_s + 2 * kStagingBytes); // 4 * f_tile floats
    Barriers* bars = reinterpret_cast<Barriers*>(reinterpret_cast<uint8_t*>(bias_s) + 1024);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int Fp = p.f_tile, KQ = p.in_q_pad >> 2;

    if (tid == 0) {
        tma_prefetch_desc(&tmx);
        tma_prefetch_desc(&tmy);
        for (int i = 0; i < kMaxXStages; ++i) {
            mbar_init(&bars->x_full[i], 1);
            mbar_init(&bars->x_empty[i], 128);
        }
        for (int i = 0; i < kASlots; ++i) {
            mbar_init(&bars->a_full[i], 128);
            mbar_init(&bars->a_empty[i], 1);
        }
        mbar_init(&bars->acc_full, 1);
        mbar_init(&bars->acc_empty, 128);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t t_acc = bars->tmem_base;
    const uint32_t t_a = t_acc + kAccCols;

    // pipeline state persists across tiles and f-tile passes
    uint32_t xs = 0, xph = 0, as = 0, aph = 0, accph = 0, sbuf = 0;

    for (int ft = 0; ft < p.n_ftiles; ++ft) {
        // ---- pack the four sub-filters of this f-tile: stored [tap][q][c*F+f] -> smem [(tap*4+c)][q/4][f][q%4], tf32-rn
        {
            const int per_tap = p.in_q_pad * 4 * Fp;
            const int total = p.taps * per_tap;
            for (int i = tid; i < total; i += kThreads) {
                const int f = i % Fp, c = (i / Fp) & 3, q = (i / (4 * Fp)) % p.in_q_pad, tap = i / per_tap;
                float v = 0.f;
                if (q < p.in_q) v = __ldg(w + ((size_t)tap * p.in_q + q) * 4 * p.F + c * p.F + ft * Fp + f);
                const uint32_t off = ((((uint32_t)(tap * 4 + c) * KQ + (q >> 2)) * Fp + f) << 4) + ((q & 3) << 2);
                *reinterpret_cast<uint32_t*>(w_s + off) = f32_to_tf32_rn(v);
            }
            for (int i = tid; i < 4 * Fp; i += kThreads)
                bias_s[i] = p.has_bias ? __ldg(bias + (i / Fp) * p.F + ft * Fp + (i % Fp)) : 0.f;
            fence_proxy_async_smem();  // generic-proxy writes above are read by the tensor core (async proxy)
            __syncthreads();
        }

        if (warp == 0) {
            // =========================== TMA producer ===========================
            if (tid == 0) {
                for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                    const int b = tile / p.tiles_per_seq, t0 = (tile % p.tiles_per_seq) * kTileM;
                    for (int a = 0; a < 4; ++a)
                        for (int ch = 0; ch < p.n_chunks; ++ch) {
                            mbar_wait(&bars->x_empty[xs], xph ^ 1);
                            mbar_arrive_expect_tx(&bars->x_full[xs], (uint32_t)p.rows_in * 128u);
                            tma_load_4d(x_s + (size_t)xs * p.x_stage_bytes, &tmx, &bars->x_full[xs], ch * 32, a,
                                        t0 - p.pad_lo, b);
                            if (++xs == (uint32_t)p.x_stages) { xs = 0; xph ^= 1; }
                        }
                }
            }
        } else if (warp == 1) {
            // =========================== MMA issuer ===========================
            if (tid == 32) {
                const uint32_t idesc_pos = idesc_tf32(kTileM, Fp, false, false);
                const uint32_t idesc_neg = idesc_tf32(kTileM, Fp, false, true);
                const uint32_t w_addr = smem_u32(w_s);
                const uint32_t lbo = (uint32_t)Fp * 16u;
                for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                    mbar_wait(&bars->acc_empty, accph ^ 1);
                    tc_fence_after_sync();
                    uint32_t accumulate = 0;
                    for (int a = 0; a < 4; ++a)
                        for (int ch = 0; ch < p.n_chunks; ++ch) {
                            const int ksteps = min(32, p.in_q_pad - ch * 32) >> 3;
                            for (int tap = 0; tap < p.taps; ++tap) {
                                mbar_wait(&bars->a_full[as], aph);
                                tc_fence_after_sync();
                                for (int ks = 0; ks < ksteps; ++ks) {
                                    const uint32_t a_col = t_a + as * kASlotCols + ks * 8;
#pragma unroll
                                    for (int b = 0; b < 4; ++b) {
                                        const int c = a ^ b;  // sub-filter index: IDX[a][b] = a xor b
                                        const uint32_t bit = p.conj_w ? (b * 4 + a) : (a * 4 + b);
                                        const uint32_t b_addr =
                                            w_addr + (((uint32_t)(tap * 4 + c) * KQ + (ch * 8 + ks * 2)) * Fp << 4);
                                        mma_tf32_ts(t_acc + b * Fp, a_col, smem_desc_kmajor_noswz(b_addr, lbo, 128),
                                                    ((kNegConv >> bit) & 1u) ? idesc_neg : idesc_pos, accumulate);
                                    }
                                    accumulate = 1;
                                }
                                mma_commit(&bars->a_empty[as]);  // slot free once these MMAs have read it
                                if (++as == kASlots) { as = 0; aph ^= 1; }
                            }
                        }
                    mma_commit(&bars->acc_full);
                    accph ^= 1;
                }
            }
        } else if (warp >= 4 && warp < 8) {
            // =========================== converters: smem fp32 -> tf32(rn) -> TMEM A slots ===========================
            const int r = tid - 128;
            const uint32_t lane_base = (uint32_t)(r & ~31) << 16;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int a = 0; a < 4; ++a)
                    for (int ch = 0; ch < p.n_chunks; ++ch) {
                        const int kc = min(32, p.in_q_pad - ch * 32);
                        mbar_wait(&bars->x_full[xs], xph);
                        const uint8_t* xb = x_s + (size_t)xs * p.x_stage_bytes;
                        for (int tap = 0; tap < p.taps; ++tap) {
                            const uint32_t row = (uint32_t)(r + tap * p.dil);
                            mbar_wait(&bars->a_empty[as], aph ^ 1);
                            tc_fence_after_sync();
                            for (int k0 = 0; k0 < kc; k0 += 8) {
                                const float4 v0 = *reinterpret_cast<const float4*>(xb + swz128(row, k0 >> 2));
                                const float4 v1 = *reinterpret_cast<const float4*>(xb + swz128(row, (k0 >> 2) + 1));
                                const uint32_t u[8] = {f32_to_tf32_rn(v0.x), f32_to_tf32_rn(v0.y), f32_to_tf32_rn(v0.z),
                                                       f32_to_tf32_rn(v0.w), f32_to_tf32_rn(v1.x), f32_to_tf32_rn(v1.y),
                                                       f32_to_tf32_rn(v1.z), f32_to_tf32_rn(v1.w)};
                                tmem_st8(t_a + lane_base + as * kASlotCols + k0, u);
                            }
                            tmem_wait_st();
                            tc_fence_before_sync();
                            mbar_arrive(&bars->a_full[as]);
                            if (++as == kASlots) { as = 0; aph ^= 1; }
                        }
                        mbar_arrive(&bars->x_empty[xs]);
                        if (++xs == (uint32_t)p.x_stages) { xs = 0; xph ^= 1; }
                    }
            }
        } else if (warp >= 8) {
            // =========================== epilogue ===========================
            const int r = tid - 256;
            const uint32_t lane_base = (uint32_t)(r & ~31) << 16;
            const int n_chunks_out = (4 * Fp) >> 5;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int b = tile / p.tiles_per_seq, t0 = (tile % p.tiles_per_seq) * kTileM;
                mbar_wait(&bars->acc_full, accph);
                tc_fence_after_sync();
                for (int c = 0; c < n_chunks_out; ++c) {
                    uint32_t v[32];
                    tmem_ld32(t_acc + lane_base + c * 32, v);
                    tmem_wait_ld();
                    if (c == n_chunks_out - 1) {  // accumulators drained: the next tile's MMAs may start
                        tc_fence_before_sync();
                        mbar_arrive(&bars->acc_empty);
                    }
                    if (r == 0) tma_store_wait_read<1>();  // the store that last used this staging tile has read it
                    epi_bar_sync();
                    uint8_t* st = y_s + sbuf * kStagingBytes;
                    const float* bs = bias_s + c * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o;
                        o.x = act_apply(__uint_as_float(v[4 * j + 0]) + bs[4 * j + 0], p.act);
                        o.y = act_apply(__uint_as_float(v[4 * j + 1]) + bs[4 * j + 1], p.act);
                        o.z = act_apply(__uint_as_float(v[4 * j + 2]) + bs[4 * j + 2], p.act);
                        o.w = act_apply(__uint_as_float(v[4 * j + 3]) + bs[4 * j + 3], p.act);
                        *reinterpret_cast<float4*>(st + swz128((uint32_t)r, (uint32_t)j)) = o;
                    }
                    fence_proxy_async_smem();
                    epi_bar_sync();
                    if (r == 0) {
                        // accumulator column c*32 -> output channel: component (c*32)/Fp, filter ft*Fp + (c*32)%Fp
                        const int col = c * 32;
                        const int ch_out = (col / Fp) * p.F + ft * Fp + (col % Fp);
                        tma_store_3d(&tmy, st, ch_out, t0, b);
                        tma_store_commit();
                    }
                    sbuf ^= 1;
                }
                accph ^= 1;
            }
            if (r == 0) tma_store_wait_all<0>();
        }
        __syncthreads();  // every role is done with this f-tile's weights
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(t_acc, 512);
}

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace

TcPlan tc_plan(const Geom& g, int rank) {
    TcPlan pl{};
    pl.ok = 0;
    auto no = [&](const char* why) {
        pl.why = why;
        return pl;
    };
    if (g.channels_first) return no("channels_first layout");
    if (rank != 1) return no("rank > 1");
    if (g.s[2] != 1) return no("stride != 1");
    if (g.in_q % 4) return no("in_q not a multiple of 4 (TMA stride alignment)");
    if (g.F % 16) return no("filters not a multiple of 16");
    int f_tile = g.F;
    if (g.F > 64) {
        f_tile = (g.F % 64 == 0) ? 64 : (g.F % 32 == 0 ? 32 : 0);
        if (!f_tile) return no("filters > 64 and not a multiple of 32");
    }
    const int taps = g.k[2];
    const int rows_in = kTileM + (taps - 1) * g.d[2];
    if (rows_in > 256) return no("halo exceeds the 256-row TMA box");
    if (g.out_sp[2] < 1 || g.batch < 1) return no("empty problem");
    const int in_q_pad = (g.in_q + 7) & ~7;
    const size_t w_bytes = (size_t)taps * 4 * in_q_pad * f_tile * 4;
    const size_t w_pad = (w_bytes + 1023) & ~size_t(1023);
    const size_t stage = ((size_t)rows_in * 128 + 1023) & ~size_t(1023);
    const size_t fixed = 1024 /*align slack*/ + w_pad + 2 * kStagingBytes + 1024 /*bias*/ + 512 /*barriers*/;
    if (fixed + 2 * stage > kSmemLimit) return no("sub-filters do not fit in shared memory");
    int stages = (int)std::min<size_t>(kMaxXStages, (kSmemLimit - fixed) / stage);
    pl.ok = 1;
    pl.f_tile = f_tile;
    pl.n_ftiles = g.F / f_tile;
    pl.in_q_pad = in_q_pad;
    pl.rows_in = rows_in;
    pl.x_stages = stages;
    pl.smem_bytes = fixed + (size_t)stages * stage;
    pl.why = "";
    return pl;
}

int tc_forward(const Geom& g, int rank, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
    const TcPlan pl = tc_plan(g, rank);
    if (!pl.ok) {
        set_error("tensor-core kernel does not take this shape: %s", pl.why);
        return QNN_E_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) {
        set_error("tensor-core kernel needs 16-byte aligned x and y");
        return QNN_E_UNSUPPORTED;
    }
    const int L = g.in_sp[2], Lo = g.out_sp[2];
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
    p.in_q = g.in_q;
    p.in_q_pad = pl.in_q_pad;
    p.n_chunks = (pl.in_q_pad + 31) / 32;
    p.F = g.F;
    p.f_tile = pl.f_tile;
    p.n_ftiles = pl.n_ftiles;
    p.rows_in = pl.rows_in;
    p.x_stages = pl.x_stages;
    p.x_stage_bytes = (int)(((size_t)pl.rows_in * 128 + 1023) & ~size_t(1023));
    p.act = g.act;
    p.conj_w = g.conj_w;
    p.has_bias = bias != nullptr;
    p.w_bytes = (uint32_t)((size_t)p.taps * 4 * p.in_q_pad * p.f_tile * 4);

    CUtensorMap tmx, tmy;
    {
        const uint64_t dims[4] = {(uint64_t)g.in_q, 4, (uint64_t)L, (uint64_t)g.batch};
        const uint64_t str[3] = {(uint64_t)g.in_q * 4, (uint64_t)g.in_q * 16, (uint64_t)L * g.in_q * 16};
        const uint32_t box[4] = {32, 1, (uint32_t)pl.rows_in, 1};
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
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(k_hamilton_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit);
    });
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
        return QNN_E_CUDA;
    }
    const int grid = std::min(p.n_tiles, num_sms());
    k_hamilton_tc<<<grid, kThreads, pl.smem_bytes, st>>>(tmx, tmy, p, w, bias);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("tensor-core kernel launch failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    return QNN_OK;
}

}  // namespace qnn
