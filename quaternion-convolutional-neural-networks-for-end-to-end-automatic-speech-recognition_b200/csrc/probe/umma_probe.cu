// Hardware bring-up probe for the building blocks the Hamilton tensor-core kernel relies on.
// Each variant is a tiny single-CTA experiment checked against a CPU loop; run as `umma_probe <variant>`.
//   0  tf32 MMA, A and B from shared memory (K-major, no swizzle)
//   1  tf32 MMA, A from tensor memory (written with tcgen05.st 32x32b), B from shared memory
//   2  as 1 with the instruction-descriptor negate-B bit     3  as 1 with the negate-A bit
//   4  does the tensor core truncate or round fp32 operands to tf32?
//   5  mini Hamilton product: 4 accumulators, 16 signed block MMAs, operands x_a in TMEM, sub-filters in smem
//   6  TMA 4-D load, 128-B swizzle, box wider than the tensor and negative start row (zero fill), thread un-swizzle
//   7  TMA 3-D store from a swizzled staging tile with clipping at the tensor edge
//   8  as 1 with N = 32 and N = 16
//  10  B operand MN-major with the TMA 128-byte swizzle (the layout a TMA box of the stored kernel lands in)
//   9  tensor-pipe rate: cycles per tf32 MMA for N = 32/64/128/256, A from TMEM or smem, 1 or 4 accumulators, 1 or 2 issuers
#define QNN_SPIN_LIMIT 20000000
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../qnn_ptx.cuh"
#include "../qnn_tmap.h"

using namespace qnn::ptx;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

// element (row,k) of a [rows x K] fp32 K-major no-swizzle operand: SBO = 128 B, LBO = rows*16 B
__device__ __forceinline__ uint32_t noswz_off(int row, int k, int rows) {
    return (row & 7) * 16 + (row >> 3) * 128 + (k >> 2) * (rows * 16) + (k & 3) * 4;
}

// D[128 x N] = (+-A)[128 x K] * (+-B)[N x K]^T
__global__ void __launch_bounds__(128) k_mma(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                             int K, int N, int a_in_tmem, int negA, int negB) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_s = smem;                  // 128*K*4
    uint8_t* b_s = smem + 128 * K * 4;    // N*K*4
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        tmem_alloc(&tmem_slot, 256);
        tmem_relinquish();
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 128 * K; i += 128) {
        int r = i / K, k = i % K;
        *reinterpret_cast<float*>(a_s + noswz_off(r, k, 128)) = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        int r = i / K, k = i % K;
        *reinterpret_cast<float*>(b_s + noswz_off(r, k, N)) = B[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t t_acc = tbase;          // columns [0, N)
    const uint32_t t_a = tbase + 128;      // columns [128, 128+K)
    const uint32_t lane_base = uint32_t(warp * 32) << 16;

    if (a_in_tmem) {
        for (int k0 = 0; k0 < K; k0 += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(A[tid * K + k0 + j]);
            tmem_st8(t_a + lane_base + k0, v);
        }
        tmem_wait_st();
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    if (tid == 0) {
        const uint32_t idesc = idesc_tf32(128, N, negA, negB);
        for (int ks = 0; ks < K / 8; ++ks) {
            uint64_t bd = smem_desc_kmajor_noswz(smem_u32(b_s) + ks * 2 * (N * 16), N * 16, 128);
            if (a_in_tmem) {
                mma_tf32_ts(t_acc, t_a + ks * 8, bd, idesc, ks > 0);
            } else {
                uint64_t ad = smem_desc_kmajor_noswz(smem_u32(a_s) + ks * 2 * (128 * 16), 128 * 16, 128);
                mma_tf32_ss(t_acc, ad, bd, idesc, ks > 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(t_acc + lane_base + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

// conv-convention Hamilton tables: y_b = sum_a S[a][b] * x_a * f_{IDX[a][b]}
__constant__ int c_IDX[4][4] = {{0, 1, 2, 3}, {1, 0, 3, 2}, {2, 3, 0, 1}, {3, 2, 1, 0}};
__constant__ int c_NEG[4][4] = {{0, 0, 0, 0}, {1, 0, 0, 1}, {1, 1, 0, 0}, {1, 0, 1, 0}};
static const int h_IDX[4][4] = {{0, 1, 2, 3}, {1, 0, 3, 2}, {2, 3, 0, 1}, {3, 2, 1, 0}};
static const int h_NEG[4][4] = {{0, 0, 0, 0}, {1, 0, 0, 1}, {1, 1, 0, 0}, {1, 0, 1, 0}};

// x [128, 4*Q] blocked, w [Q, 4*F] blocked, y [128, 4*F]; Q = 8, F = 64
__global__ void __launch_bounds__(128) k_hamilton(const float* __restrict__ x, const float* __restrict__ w,
                                                  float* __restrict__ y) {
    constexpr int Q = 8, F = 64;
    extern __shared__ __align__(1024) uint8_t smem[];  // 4 sub-filters, each [F rows x Q] K-major no-swizzle
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        tmem_alloc(&tmem_slot, 512);
        tmem_relinquish();
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < Q * 4 * F; i += 128) {
        int q = i / (4 * F), n = i % (4 * F), c = n / F, f = n % F;
        *reinterpret_cast<float*>(smem + c * (F * Q * 4) + noswz_off(f, q, F)) = w[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t t_a = tbase + 256;
    const uint32_t lane_base = uint32_t(warp * 32) << 16;
    for (int a = 0; a < 4; ++a) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(x[tid * 4 * Q + a * Q + j]);
        tmem_st8(t_a + lane_base + a * 8, v);
    }
    tmem_wait_st();
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (tid == 0) {
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) {
                uint64_t bd = smem_desc_kmajor_noswz(smem_u32(smem) + c_IDX[a][b] * (F * Q * 4), F * 16, 128);
                mma_tf32_ts(tbase + b * F, t_a + a * 8, bd, idesc_tf32(128, F, false, c_NEG[a][b] != 0), a > 0);
            }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < 4 * F; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(tbase + lane_base + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) y[tid * 4 * F + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// TMA load probe: one box -> smem (swizzled) -> un-swizzled copy to out[rows][32]
__global__ void __launch_bounds__(128) k_tma_load(const __grid_constant__ CUtensorMap tmap, float* out, int rows, int c0,
                                                  int c1, int c2, int c3) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, rows * 128);
        tma_load_4d(smem, &tmap, &bar, c0, c1, c2, c3);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < rows * 32; i += 128) {
        int r = i / 32, c = i % 32;
        out[i] = *reinterpret_cast<float*>(smem + swz128(r, c >> 2) + (c & 3) * 4);
    }
}

// TMA store probe: staging[rows][32] written swizzled by threads, stored with a 3-D map
__global__ void __launch_bounds__(128) k_tma_store(const __grid_constant__ CUtensorMap tmap, int rows, int c0, int c1,
                                                   int c2) {
    extern __shared__ __align__(1024) uint8_t smem[];
    for (int i = threadIdx.x; i < rows * 32; i += 128) {
        int r = i / 32, c = i % 32;
        *reinterpret_cast<float*>(smem + swz128(r, c >> 2) + (c & 3) * 4) = 1000.f * r + c;
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_store_3d(&tmap, smem, c0, c1, c2);
        tma_store_commit();
        tma_store_wait_all<0>();
    }
}

// `n_issuers` threads (lane 0 of warps 0..n_issuers-1) each issue reps / n_issuers MMAs (K = 8 each) into their own
// accumulator columns; reports the SM cycles from first issue to completion of everything.
__global__ void __launch_bounds__(256) k_rate(int N, int a_in_tmem, int n_issuers, int same_acc, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        tmem_alloc(&tmem_slot, 512);
        tmem_relinquish();
    }
    if (tid == 0) {
        mbar_init(&bar, n_issuers);
        fence_mbar_init();
    }
    for (int i = tid; i < (128 + 256) * 32 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    long long t0 = clock64();
    if ((tid & 31) == 0 && warp < n_issuers) {
        const uint32_t idesc = idesc_tf32(128, N, false, false);
        const uint64_t bd = smem_desc_kmajor_noswz(smem_u32(smem) + 128 * 32, N * 16, 128);
        const uint64_t ad = smem_desc_kmajor_noswz(smem_u32(smem), 128 * 16, 128);
        const int my_reps = reps / n_issuers;
        // accumulator columns: issuers share 256 columns; with same_acc every MMA of a thread hits one accumulator,
        // otherwise it alternates between two
        const uint32_t d0 = tbase + (uint32_t)((warp * 2 * N) % 256);
        const uint32_t d1 = same_acc ? d0 : tbase + (uint32_t)(((warp * 2 + 1) * N) % 256);
        const uint32_t ta = tbase + 256;
        for (int i = 0; i < my_reps; i += 4) {
            if (a_in_tmem) {
                mma_tf32_ts(d0, ta, bd, idesc, 1);
                mma_tf32_ts(d1, ta + 8, bd, idesc, 1);
                mma_tf32_ts(d0, ta + 16, bd, idesc, 1);
                mma_tf32_ts(d1, ta + 24, bd, idesc, 1);
            } else {
                mma_tf32_ss(d0, ad, bd, idesc, 1);
                mma_tf32_ss(d1, ad, bd, idesc, 1);
                mma_tf32_ss(d0, ad, bd, idesc, 1);
                mma_tf32_ss(d1, ad, bd, idesc, 1);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (tid == 0) out[blockIdx.x] = clock64() - t0;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// D[128 x N] = A[128 x K] * Bkn[K x N]; Bkn is row-major [k][n] in global (n contiguous, like the stored kernel).
// smem image = what TMA SWIZZLE_128B would write for boxes of 32 n by K rows: panel h (n in [32h, 32h+32)) at
// b_s + h*K*128, row k at +k*128, 16-byte chunk c at ((c ^ (k & 7)) << 4).
__global__ void __launch_bounds__(128) k_mma_bmn(const float* __restrict__ A, const float* __restrict__ Bkn,
                                                 float* __restrict__ D, int K, int N, int a_in_tmem) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_s = smem;                       // 128*K*4 (K-major no swizzle)
    uint8_t* b_s = smem + ((128 * K * 4 + 1023) & ~1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        tmem_alloc(&tmem_slot, 512);
        tmem_relinquish();
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 128 * K; i += 128) {
        int r = i / K, k = i % K;
        *reinterpret_cast<float*>(a_s + noswz_off(r, k, 128)) = A[i];
    }
    for (int i = tid; i < K * N; i += 128) {
        int k = i / N, n = i % N, h = n / 32, c = (n % 32) / 4, w4 = n % 4;
        *reinterpret_cast<float*>(b_s + h * K * 128 + k * 128 + ((c ^ (k & 7)) << 4) + w4 * 4) = Bkn[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t t_acc = tbase, t_a = tbase + 256;
    const uint32_t lane_base = uint32_t(warp * 32) << 16;
    if (a_in_tmem) {
        for (int k0 = 0; k0 < K; k0 += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(A[tid * K + k0 + j]);
            tmem_st8(t_a + lane_base + k0, v);
        }
        tmem_wait_st();
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = idesc_tf32_b_mn(128, N, false, false);
        for (int ks = 0; ks < K / 8; ++ks) {
            // K step ks = rows 8ks..8ks+7 of every panel: +ks*1024 bytes; panels are K*128 bytes apart
            uint64_t bd = smem_desc_mnmajor_sw128(smem_u32(b_s) + ks * 1024, K * 128, 1024);
            if (a_in_tmem) {
                mma_tf32_ts(t_acc, t_a + ks * 8, bd, idesc, ks > 0);
            } else {
                uint64_t ad = smem_desc_kmajor_noswz(smem_u32(a_s) + ks * 2 * (128 * 16), 128 * 16, 128);
                mma_tf32_ss(t_acc, ad, bd, idesc, ks > 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(t_acc + lane_base + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

static int run_mma_bmn(int K, int N, int a_tmem) {
    std::vector<float> A(128 * K), B(K * N), D(128 * N), R(128 * N);
    for (auto& v : A) v = float((rand() % 9) - 4);
    for (auto& v : B) v = float((rand() % 9) - 4);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < K; ++k) s += A[m * K + k] * B[k * N + n];
            R[m * N + n] = s;
        }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4));
    CK(cudaMalloc(&dB, B.size() * 4));
    CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, D.size() * 4));
    k_mma_bmn<<<1, 128, 128 * K * 4 + 1024 + K * N * 4 + 1024>>>(dA, dB, dD, K, N, a_tmem);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (size_t i = 0; i < D.size(); ++i) bad += D[i] != R[i];
    printf("v10 B MN-major SW128: K=%d N=%d a_tmem=%d bad=%d  D[0..3]=%g %g %g %g ref %g %g %g %g -> %s\n", K, N, a_tmem,
           bad, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3], bad ? "FAIL" : "PASS");
    return bad != 0;
}

static float frand_int() { return float((rand() % 9) - 4); }

static int run_mma(int K, int N, int a_tmem, int negA, int negB, const char* name) {
    std::vector<float> A(128 * K), B(N * K), D(128 * N), R(128 * N);
    for (auto& v : A) v = frand_int();
    for (auto& v : B) v = frand_int();
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
            R[m * N + n] = (negA != negB) ? -s : s;
        }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4));
    CK(cudaMalloc(&dB, B.size() * 4));
    CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, D.size() * 4));
    size_t sm = (128 + N) * K * 4;
    k_mma<<<1, 128, sm>>>(dA, dB, dD, K, N, a_tmem, negA, negB);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int bad = 0;
    for (size_t i = 0; i < D.size(); ++i) {
        double e = fabs((double)D[i] - R[i]);
        if (!(e <= 0)) ++bad;
        if (e > maxerr || e != e) maxerr = e;
    }
    printf("%s: K=%d N=%d a_tmem=%d negA=%d negB=%d  maxerr=%g bad=%d  D[0..3]=%g %g %g %g ref %g %g %g %g -> %s\n", name,
           K, N, a_tmem, negA, negB, maxerr, bad, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3], bad ? "FAIL" : "PASS");
    return bad != 0;
}

int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    srand(1234 + variant);
    CK(cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    int rc = 0;
    switch (variant) {
        case 0: rc = run_mma(32, 64, 0, 0, 0, "v0 SS"); break;
        case 1: rc = run_mma(32, 64, 1, 0, 0, "v1 TS"); break;
        case 2: rc = run_mma(32, 64, 1, 0, 1, "v2 TS negB"); break;
        case 3: rc = run_mma(32, 64, 1, 1, 0, "v3 TS negA"); break;
        case 8:
            rc = run_mma(32, 32, 1, 0, 0, "v8 TS N32");
            rc |= run_mma(32, 16, 1, 0, 1, "v8 TS N16 negB");
            rc |= run_mma(8, 64, 1, 0, 0, "v8 TS K8");
            rc |= run_mma(32, 128, 0, 0, 1, "v8 SS N128 negB");
            break;
        case 4: {
            // A[m][0] = 1 + (m%8) * 2^-13  (tf32 ulp at 1.0 is 2^-10); B[n][0] = 1, everything else 0.
            const int K = 8, N = 64;
            std::vector<float> A(128 * K, 0.f), B(N * K, 0.f), D(128 * N);
            for (int m = 0; m < 128; ++m) A[m * K] = 1.0f + float(m % 8) * ldexpf(1.f, -13);
            for (int n = 0; n < N; ++n) B[n * K] = 1.0f + float(n % 8) * ldexpf(1.f, -13);
            float *dA, *dB, *dD;
            CK(cudaMalloc(&dA, A.size() * 4));
            CK(cudaMalloc(&dB, B.size() * 4));
            CK(cudaMalloc(&dD, D.size() * 4));
            CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
            for (int a_tmem = 0; a_tmem < 2; ++a_tmem) {
                k_mma<<<1, 128, (128 + N) * K * 4>>>(dA, dB, dD, K, N, a_tmem, 0, 0);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
                printf("v4 rounding probe a_tmem=%d: A row m has 1+m*2^-13, B col 0 is 1. D[m][0]-1 in units of 2^-13:", a_tmem);
                for (int m = 0; m < 8; ++m) printf(" %g", (D[m * N] - 1.0f) * 8192.f);
                printf("   | B side (row 0, col n):");
                for (int n = 0; n < 8; ++n) printf(" %g", (D[n] - 1.0f) * 8192.f);
                printf("\n   (truncation -> all 0;  round-to-nearest -> 0 0 0 0 8 8 8 8 or ties variant)\n");
            }
            break;
        }
        case 5: {
            constexpr int Q = 8, F = 64;
            std::vector<float> x(128 * 4 * Q), w(Q * 4 * F), y(128 * 4 * F), r(128 * 4 * F, 0.f);
            for (auto& v : x) v = frand_int();
            for (auto& v : w) v = frand_int();
            for (int m = 0; m < 128; ++m)
                for (int b = 0; b < 4; ++b)
                    for (int f = 0; f < F; ++f) {
                        float s = 0;
                        for (int a = 0; a < 4; ++a)
                            for (int q = 0; q < Q; ++q) {
                                float p = x[m * 4 * Q + a * Q + q] * w[q * 4 * F + h_IDX[a][b] * F + f];
                                s += h_NEG[a][b] ? -p : p;
                            }
                        r[m * 4 * F + b * F + f] = s;
                    }
            float *dx, *dw, *dy;
            CK(cudaMalloc(&dx, x.size() * 4));
            CK(cudaMalloc(&dw, w.size() * 4));
            CK(cudaMalloc(&dy, y.size() * 4));
            CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemset(dy, 0xff, y.size() * 4));
            k_hamilton<<<1, 128, 4 * F * Q * 4>>>(dx, dw, dy);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(y.data(), dy, y.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0, badb[4] = {0, 0, 0, 0};
            for (size_t i = 0; i < y.size(); ++i)
                if (y[i] != r[i]) {
                    ++bad;
                    ++badb[(i % (4 * F)) / F];
                }
            printf("v5 hamilton 16-block TS MMA with negate bits: bad=%d (per output block %d %d %d %d) y[0]=%g ref %g -> %s\n",
                   bad, badb[0], badb[1], badb[2], badb[3], y[0], r[0], bad ? "FAIL" : "PASS");
            rc = bad != 0;
            break;
        }
        case 6: {
            // x[B=2][T=20][4*Q] with Q = 8 (box of 32 channels is wider than Q) and Q = 40 (second chunk half out of range)
            for (int Q : {8, 40}) {
                const int Bn = 2, T = 20, C = 4 * Q, rows = 16;
                std::vector<float> x(Bn * T * C);
                for (size_t i = 0; i < x.size(); ++i) x[i] = float(i % 4093) + 1.f;
                float *dx, *dout;
                CK(cudaMalloc(&dx, x.size() * 4));
                CK(cudaMalloc(&dout, rows * 32 * 4));
                CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
                CUtensorMap tm;
                uint64_t dims[4] = {(uint64_t)Q, 4, (uint64_t)T, (uint64_t)Bn};
                uint64_t str[3] = {(uint64_t)Q * 4, (uint64_t)C * 4, (uint64_t)T * C * 4};
                uint32_t box[4] = {32, 1, (uint32_t)rows, 1};
                int e = qnn::make_tmap_f32(&tm, dx, 4, dims, str, box, true);
                if (e) {
                    printf("v6 Q=%d: tensor map encode failed %d -> FAIL\n", Q, e);
                    rc = 1;
                    continue;
                }
                CK(cudaFuncSetAttribute(k_tma_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
                struct Case { int c0, a, t0, b; } cases[3] = {{0, 2, -1, 1}, {32, 1, 10, 0}, {0, 3, 12, 1}};
                for (auto cs : cases) {
                    if (cs.c0 >= Q) continue;
                    CK(cudaMemset(dout, 0xff, rows * 32 * 4));
                    k_tma_load<<<1, 128, rows * 128 + 1024>>>(tm, dout, rows, cs.c0, cs.a, cs.t0, cs.b);
                    CK(cudaDeviceSynchronize());
                    std::vector<float> o(rows * 32);
                    CK(cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost));
                    int bad = 0;
                    for (int r = 0; r < rows; ++r)
                        for (int c = 0; c < 32; ++c) {
                            int t = cs.t0 + r, q = cs.c0 + c;
                            float ref = (t < 0 || t >= T || q >= Q) ? 0.f : x[(cs.b * T + t) * C + cs.a * Q + q];
                            if (o[r * 32 + c] != ref) ++bad;
                        }
                    printf("v6 TMA 4D load Q=%d coords(c0=%d a=%d t0=%d b=%d): bad=%d -> %s\n", Q, cs.c0, cs.a, cs.t0, cs.b,
                           bad, bad ? "FAIL" : "PASS");
                    rc |= bad != 0;
                }
            }
            break;
        }
        case 7: {
            const int Bn = 2, T = 20, C = 96, rows = 16;
            std::vector<float> y(Bn * T * C, -1.f);
            float* dy;
            CK(cudaMalloc(&dy, y.size() * 4));
            CK(cudaMemcpy(dy, y.data(), y.size() * 4, cudaMemcpyHostToDevice));
            CUtensorMap tm;
            uint64_t dims[3] = {(uint64_t)C, (uint64_t)T, (uint64_t)Bn};
            uint64_t str[2] = {(uint64_t)C * 4, (uint64_t)T * C * 4};
            uint32_t box[3] = {32, (uint32_t)rows, 1};
            int e = qnn::make_tmap_f32(&tm, dy, 3, dims, str, box, true);
            if (e) {
                printf("v7: tensor map encode failed %d -> FAIL\n", e);
                return 1;
            }
            CK(cudaFuncSetAttribute(k_tma_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
            k_tma_store<<<1, 128, rows * 128 + 1024>>>(tm, rows, 32, 8, 1);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(y.data(), dy, y.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int b = 0; b < Bn; ++b)
                for (int t = 0; t < T; ++t)
                    for (int c = 0; c < C; ++c) {
                        float ref = -1.f;
                        if (b == 1 && t >= 8 && t < 8 + rows && c >= 32 && c < 64) ref = 1000.f * (t - 8) + (c - 32);
                        if (y[(b * T + t) * C + c] != ref) ++bad;
                    }
            printf("v7 TMA 3D store (swizzled staging, clipped at T): bad=%d -> %s\n", bad, bad ? "FAIL" : "PASS");
            rc = bad != 0;
            break;
        }
        case 10:
            CK(cudaFuncSetAttribute(k_mma_bmn, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            rc = run_mma_bmn(8, 32, 1);
            rc |= run_mma_bmn(8, 64, 1);
            rc |= run_mma_bmn(40, 64, 1);
            rc |= run_mma_bmn(40, 64, 0);
            rc |= run_mma_bmn(16, 128, 1);
            break;
        case 9: {
            CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            long long* dout;
            CK(cudaMalloc(&dout, 148 * 8));
            const int reps = 4096;
            printf("v9 tensor-pipe rate (tf32, M=128, K=8 per MMA, %d MMAs in total): cycles per MMA, aggregate  [pipe floor N/2]\n", reps);
            for (int N : {32, 64, 128, 256})
                for (int ts = 0; ts < 2; ++ts)
                    for (int same = 0; same < 2; ++same)
                        for (int ni : {1, 2, 4, 8}) {
                            if (N == 256 && !same) continue;
                            if (N == 128 && ni > 2 && !same) continue;
                            k_rate<<<148, 256, 48 * 1024>>>(N, ts, ni, same, reps, dout);
                            CK(cudaDeviceSynchronize());
                            long long h[148];
                            CK(cudaMemcpy(h, dout, 148 * 8, cudaMemcpyDeviceToHost));
                            long long mx = 0;
                            for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
                            printf("  N=%3d A=%s %s issuers=%d : %.1f cyc/MMA\n", N, ts ? "tmem" : "smem",
                                   same ? "1 acc/thread " : "2 accs/thread", ni, (double)mx / reps);
                        }
            break;
        }
        default: printf("unknown variant\n"); rc = 3;
    }
    return rc;
}
