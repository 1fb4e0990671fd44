// Hardware probe: what does a TMA stage load of a RAGGED channels_last row cost?  x[65536 rows][4 * 41 floats] (the TIMIT
// first layer of the cfg 3 stack; 656-byte rows) against x[65536][4 * 40] (cfg 2, 640-byte rows).  One CTA per SM walks
// its tiles of 128 (+2 halo) rows like k_hamilton_tc and pulls each tile's stages through a ring of `depth` buffers; one
// thread issues, waits for the oldest stage and re-uses its buffer (no consumer work).  Reported: time per pass over the
// tensor, cycles per box per SM.
//   mode 0  cfg 2: 5 boxes [32 ch x 130 rows], SWIZZLE_128B, 128-byte aligned starts
//   mode 1  ragged as shipped: 8 boxes [36 ch x 130 rows] un-swizzled, start = (41 a + 32 c) & ~3 (16-byte aligned)
//   mode 2  8 boxes [40 ch x 130 rows] un-swizzled, start = (41 a + 32 c) & ~7 (32-byte = sector aligned)
//   mode 3  12 boxes [20 ch x 130 rows] un-swizzled, start = (41 a + 16 c) & ~3
//   mode 4  whole rows: 5 boxes [164 ch x 26 rows] un-swizzled (contiguous 17 KB each)
//   mode 5  whole rows as 1-D bulk copies: 5 copies of 26 rows x 656 B (contiguous, no tensor map)
//   mode 6  cfg 2 as ONE 3-D box per tile: (32 ch, 5 chunks, 130 rows) SWIZZLE_128B = 83 KB
//   mode 7  cfg 2 as 3-D boxes of (32 ch, 1 chunk, 130 rows) through the same 3-D map (5 per tile; control for mode 6)
//   mode 8  ragged whole rows as ONE box per tile: [164 ch x 130 rows] un-swizzled = 85 KB
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_ragged_probe tma_ragged_probe.cu && ./tma_ragged_probe
#include <cstdio>
#include <cstdlib>
#include "../qnn_tmap.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kRows = 65536, kTile = 128, kTiles = kRows / kTile, kHalo = 130;
constexpr int kSlot = 88 * 1024;  // bytes per ring buffer (largest box: 130 x 656 B = 85 KB); big boxes run at depth 2

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Maps {
    CUtensorMap m[8];
};

__global__ void __launch_bounds__(128, 1) k_load(const __grid_constant__ Maps maps, const float* x41, int mode, int depth, int reps,
                                                 unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int n_box = mode == 0 ? 5 : mode == 1 ? 8 : mode == 2 ? 8 : mode == 3 ? 12 : mode == 6 ? 1 : mode == 8 ? 1 : 5;
    const int my_tiles = (kTiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const long long total = (long long)my_tiles * n_box * reps;
    auto issue = [&](long long i) {
        const int slot = (int)(i % depth);
        const long long bi = i % ((long long)my_tiles * n_box);
        const int tile = (int)blockIdx.x + (int)(bi / n_box) * (int)gridDim.x, s = (int)(bi % n_box);
        const int row0 = tile * kTile - 1;
        uint32_t bytes;
        int c0, r0 = row0;
        const CUtensorMap* tm = &maps.m[mode == 5 ? 4 : mode == 8 ? 5 : mode];
        if (mode == 0) { c0 = s * 32; bytes = kHalo * 128; }
        else if (mode == 1) { c0 = ((s / 2) * 41 + (s % 2) * 32) & ~3; bytes = kHalo * 144; }
        else if (mode == 2) { c0 = ((s / 2) * 41 + (s % 2) * 32) & ~7; bytes = kHalo * 160; }
        else if (mode == 3) { c0 = ((s / 3) * 41 + (s % 3) * 16) & ~3; bytes = kHalo * 80; }
        else if (mode == 6) { c0 = 0; bytes = kHalo * 640; }
        else if (mode == 7) { c0 = s; bytes = kHalo * 128; }
        else if (mode == 8) { c0 = 0; bytes = kHalo * 656; }
        else { c0 = 0; r0 = row0 + s * 26; bytes = 26 * 656; }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[slot])), "r"(bytes) : "memory");
        if (mode == 6 || mode == 7) {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
                         "r"(smem_u32(smem + (size_t)slot * kSlot)), "l"(tm), "r"(smem_u32(&bar[slot])), "r"(0), "r"(c0), "r"(r0) : "memory");
        } else if (mode != 5) {
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
                         "r"(smem_u32(smem + (size_t)slot * kSlot)), "l"(tm), "r"(smem_u32(&bar[slot])), "r"(c0), "r"(r0) : "memory");
        } else {
            const int rr = r0 < 0 ? 0 : r0;  // (no zero fill: clamp the first row)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                         "r"(smem_u32(smem + (size_t)slot * kSlot)), "l"(x41 + (size_t)rr * 164), "r"(bytes), "r"(smem_u32(&bar[slot])) : "memory");
        }
    };
    const unsigned long long t0 = clock64();
    long long issued = 0;
    for (; issued < depth && issued < total; ++issued) issue(issued);
    uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long done = 0; done < total; ++done) {
        const int slot = (int)(done % depth);
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[slot])), "r"(ph[slot]) : "memory");
        ph[slot] ^= 1;
        if (issued < total) issue(issued++);
    }
    cycles[blockIdx.x] = clock64() - t0;
}

int main() {
    float *x40, *x41;
    CK(cudaMalloc(&x40, (size_t)kRows * 160 * 4));
    CK(cudaMalloc(&x41, (size_t)(kRows + 64) * 164 * 4));
    CK(cudaMemset(x40, 0, (size_t)kRows * 160 * 4));
    CK(cudaMemset(x41, 0, (size_t)(kRows + 64) * 164 * 4));
    unsigned long long* cyc;
    CK(cudaMallocManaged(&cyc, 148 * 8));
    Maps maps;
    {
        const uint64_t d40[2] = {160, kRows}, s40[1] = {640};
        const uint64_t d41[2] = {164, kRows}, s41[1] = {656};
        const uint32_t b0[2] = {32, kHalo}, b1[2] = {36, kHalo}, b2[2] = {40, kHalo}, b3[2] = {20, kHalo}, b4[2] = {164, 26};
        int e = qnn::make_tmap_f32(&maps.m[0], x40, 2, d40, s40, b0, true);
        e |= qnn::make_tmap_f32(&maps.m[1], x41, 2, d41, s41, b1, false);
        e |= qnn::make_tmap_f32(&maps.m[2], x41, 2, d41, s41, b2, false);
        e |= qnn::make_tmap_f32(&maps.m[3], x41, 2, d41, s41, b3, false);
        e |= qnn::make_tmap_f32(&maps.m[4], x41, 2, d41, s41, b4, false);
        {
            const uint64_t d3[3] = {32, 5, kRows}, s3[2] = {128, 640};
            const uint32_t b6[3] = {32, 5, kHalo}, b7[3] = {32, 1, kHalo}, b8[2] = {164, kHalo};
            e |= qnn::make_tmap_f32(&maps.m[6], x40, 3, d3, s3, b6, true);
            e |= qnn::make_tmap_f32(&maps.m[7], x40, 3, d3, s3, b7, true);
            e |= qnn::make_tmap_f32(&maps.m[8 - 1 + 0], x40, 3, d3, s3, b7, true);  // (slot 7 again: keeps the array dense)
            e |= qnn::make_tmap_f32(&maps.m[5], x41, 2, d41, s41, b8, false);       // mode 8 reads m[5] below
        }
        if (e) {
            printf("tensor map encoding failed (%d)\n", e);
            return 1;
        }
    }
    const int smem = 1024 + 2 * kSlot;
    CK(cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const char* names[9] = {"cfg 2: 5 x [32 ch] swizzled, 128 B aligned", "ragged: 8 x [36 ch] 16 B aligned", "ragged: 8 x [40 ch] 32 B aligned",
                            "ragged: 12 x [20 ch] 16 B aligned", "whole rows: 5 x [164 ch x 26 rows] boxes", "whole rows: 5 x 17 KB bulk copies",
                            "cfg 2: ONE 3-D box (32, 5, 130) swizzled 83 KB", "cfg 2: 5 x 3-D box (32, 1, 130) swizzled", "ragged: ONE box [164 ch x 130 rows] 85 KB"};
    const int nbox[9] = {5, 8, 8, 12, 5, 5, 1, 5, 1};
    for (int depth : {1, 2}) {
        for (int mode = 0; mode < 9; ++mode) {
            const int reps = 6;
            k_load<<<148, 128, smem>>>(maps, x41, mode, depth, 1, cyc);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            k_load<<<148, 128, smem>>>(maps, x41, mode, depth, reps, cyc);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            unsigned long long mx = 0;
            for (int i = 0; i < 148; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
            const double boxes = 4.0 * nbox[mode] * reps;  // of a 4-tile CTA
            printf("depth %d mode %d  %-46s : %7.2f us per pass, %6.0f cycles per box per SM, %6.0f cycles per tile\n", depth, mode, names[mode],
                   ms * 1e3 / reps, (double)mx / boxes, (double)mx / (4.0 * reps));
        }
    }
    return 0;
}
