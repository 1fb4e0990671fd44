// Hardware probe: how fast can the SMs WRITE an output tensor shaped like the hot path's y[rows][256] fp32 (row pitch 1 KB)?
// One CTA per SM writes tiles of 128 rows x 256 channels (131 KB) out of shared memory, tile t by CTA t % grid -- the
// epilogue traffic of k_hamilton_tc without anything else running.  Modes:
//   0  st.global.v4 from registers, a warp instruction = 512 contiguous bytes (4 rows x 128 B chunk column)
//   1  TMA tensor store, box [32 ch x 128 rows], SWIZZLE_128B  (what the kernel does: 8 stores per tile)
//   2  TMA tensor store, box [256 ch x 16 rows], no swizzle     (whole 1 KB rows: 8 stores per tile)
//   3  1-D bulk store, one 1 KB row per copy (128 copies per tile, one per thread)
//   4  1-D bulk store, 16 KB contiguous per copy (8 per tile; upper bound: the tile as one contiguous block)
//   5  TMA tensor store, box [64 ch x 64 rows], no swizzle      (256-byte row segments: 8 stores per tile)
//   6  mode 1 with a read stream running beside it: every CTA also TMA-loads 83 KB per tile (the x traffic of a dense tile)
// Prints GB/s of written bytes per mode, for the full grid and for 68 CTAs (the last round of cfg 2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o store_probe store_probe.cu && ./store_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../qnn_tmap.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kRows = 65536, kCh = 256, kTile = 128;
constexpr int kTiles = kRows / kTile;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512, 1)
k_store(const __grid_constant__ CUtensorMap tm32, const __grid_constant__ CUtensorMap tm256, const __grid_constant__ CUtensorMap tm64,
        const __grid_constant__ CUtensorMap tmx, float* y, const float* x, int mode, int reps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    for (int i = tid; i < 128 * 1024 / 4; i += 512) reinterpret_cast<float*>(smem)[i] = (float)i;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint32_t ph = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int tile = blockIdx.x; tile < kTiles; tile += gridDim.x) {
            const int r0 = tile * kTile;
            if (mode == 0) {
                // thread -> (row, 16-byte unit): consecutive lanes walk a 128-byte chunk column, 4 rows per warp instruction
                const float4* s4 = reinterpret_cast<const float4*>(smem);
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int idx = i * 512 + tid;             // 8192 float4 = 128 rows x 64 units
                    const int chunk = idx >> 10, rem = idx & 1023, row = rem >> 3, u = rem & 7;
                    *reinterpret_cast<float4*>(y + (size_t)(r0 + row) * kCh + chunk * 32 + u * 4) = s4[idx];
                }
            } else if (mode == 1 || mode == 6) {
                if (mode == 6 && tid == 32) {
                    // the read stream: 5 boxes of [32 ch x 128 rows] (16 KB each) into the upper part of shared memory
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(5 * 16384) : "memory");
                    for (int c = 0; c < 5; ++c)
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
                                     "r"(smem_u32(smem + 131072 + c * 16384)), "l"(&tmx), "r"(smem_u32(&bar)), "r"(c * 32), "r"(r0 + (rep % 3) * kRows) : "memory");
                }
                if (tid == 0) {
                    for (int c = 0; c < 8; ++c)
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm32),
                                     "r"(smem_u32(smem + c * 16384)), "r"(c * 32), "r"(r0) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                if (mode == 6) {
                    uint32_t done = 0;
                    while (!done)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(ph) : "memory");
                    ph ^= 1;
                }
            } else if (mode == 2) {
                if (tid == 0) {
                    for (int c = 0; c < 8; ++c)
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm256),
                                     "r"(smem_u32(smem + c * 16384)), "r"(0), "r"(r0 + c * 16) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
            } else if (mode == 5) {
                if (tid == 0) {
                    for (int c = 0; c < 8; ++c)
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm64),
                                     "r"(smem_u32(smem + c * 16384)), "r"((c & 3) * 64), "r"(r0 + (c >> 2) * 64) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
            } else if (mode == 3) {
                if (tid < 128) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(y + (size_t)(r0 + tid) * kCh),
                                 "r"(smem_u32(smem + tid * 1024)), "r"(1024) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
            } else if (mode == 4) {
                if (tid < 8) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(y + (size_t)r0 * kCh + tid * 4096),
                                 "r"(smem_u32(smem + tid * 16384)), "r"(16384) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
            }
            __syncthreads();
        }
    }
    if (tid < 128) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    float *y, *x;
    CK(cudaMalloc(&y, (size_t)kRows * kCh * 4));
    CK(cudaMalloc(&x, (size_t)kRows * 160 * 4 * 3));
    CK(cudaMemset(y, 0, (size_t)kRows * kCh * 4));
    CK(cudaMemset(x, 0, (size_t)kRows * 160 * 4 * 3));
    CUtensorMap tm32, tm256, tm64, tmx;
    {
        const uint64_t dims[2] = {kCh, kRows};
        const uint64_t str[1] = {kCh * 4};
        const uint32_t b32[2] = {32, 128}, b256[2] = {256, 16}, b64[2] = {64, 64};
        if (qnn::make_tmap_f32(&tm32, y, 2, dims, str, b32, true) || qnn::make_tmap_f32(&tm256, y, 2, dims, str, b256, false) ||
            qnn::make_tmap_f32(&tm64, y, 2, dims, str, b64, false)) {
            printf("tensor map encoding failed\n");
            return 1;
        }
        const uint64_t xd[2] = {160, 3 * kRows};  // three input sets, rotated per pass (126 MB: misses in L2 like the bench)
        const uint64_t xs[1] = {160 * 4};
        const uint32_t xb[2] = {32, 128};
        if (qnn::make_tmap_f32(&tmx, x, 2, xd, xs, xb, true)) {
            printf("tensor map encoding failed (x)\n");
            return 1;
        }
    }
    const int smem = 1024 + 131072 + 5 * 16384;
    CK(cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const char* names[7] = {"st.global.v4 (512 B per warp instruction)", "TMA store [32 ch x 128 rows] swizzle 128B", "TMA store [256 ch x 16 rows]",
                            "bulk store 1 KB rows", "bulk store 16 KB contiguous", "TMA store [64 ch x 64 rows]",
                            "TMA store [32 x 128] + TMA loads 83 KB per tile"};
    for (int grid : {148, 68}) {
        for (int mode = 0; mode < 7; ++mode) {
            const int reps = 20;
            k_store<<<grid, 512, smem>>>(tm32, tm256, tm64, tmx, y, x, mode, 2);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            k_store<<<grid, 512, smem>>>(tm32, tm256, tm64, tmx, y, x, mode, reps);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double bytes = (double)kRows * kCh * 4 * reps;
            printf("grid %3d mode %d  %-50s : %7.1f GB/s written, %6.2f us per 67 MB pass, %5.1f B/clk/SM @1.9 GHz\n", grid, mode, names[mode],
                   bytes / ms / 1e6, ms * 1e3 / reps, bytes / ms / 1e6 / grid / 1.9);
        }
    }
    return 0;
}
