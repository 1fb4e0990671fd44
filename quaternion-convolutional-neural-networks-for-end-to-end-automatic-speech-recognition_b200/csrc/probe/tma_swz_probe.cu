// Probe: where does a 5-D TMA box with a 32-byte inner dimension land in shared memory under SWIZZLE_128B?
// Tensor x[N][H][W][4][Q] (channels_last, component-blocked), box (8 q, 4 components, bw positions, 1, 1).
// usage: tma_swz_probe Q W H bw x0   -> prints, per 16-byte chunk of shared memory, which (w, comp, q) it holds, and checks
// the hypothesis "dense box order [w][comp][q], 16-byte chunk index XORed with (w & 7)".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../qnn_ptx.cuh"
#include "../qnn_tmap.h"
using namespace qnn;
using namespace qnn::ptx;

__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int n, int x0) {
    extern __shared__ uint8_t sm_raw[];
    uint8_t* sm = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < n; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = -1.f;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_proxy_async_smem();
        mbar_arrive_expect_tx(&bar, n * 4);
        tma_load_5d(sm, &tm, &bar, 0, 0, x0, 1, 0);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}

int main(int argc, char** argv) {
    int Q = argc > 1 ? atoi(argv[1]) : 8, W = argc > 2 ? atoi(argv[2]) : 40, H = argc > 3 ? atoi(argv[3]) : 3;
    int bw = argc > 4 ? atoi(argv[4]) : 20, x0 = argc > 5 ? atoi(argv[5]) : -1;
    const int N = 1;
    size_t total = (size_t)N * H * W * 4 * Q;
    std::vector<float> h(total);
    for (size_t i = 0; i < total; ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, total * 4);
    cudaMemcpy(d, h.data(), total * 4, cudaMemcpyHostToDevice);
    int n = 32 * bw;
    cudaMalloc(&o, n * 4);
    CUtensorMap tm;
    uint64_t dims[5] = {(uint64_t)Q, 4, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[4] = {(uint64_t)Q * 4, (uint64_t)Q * 16, (uint64_t)W * Q * 16, (uint64_t)H * W * Q * 16};
    uint32_t box[5] = {8, 4, (uint32_t)bw, 1, 1};
    int e = make_tmap_f32(&tm, d, 5, dims, str, box, true);
    if (e) { printf("encode failed %d\n", e); return 1; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<<<1, 128, 100 * 1024>>>(tm, o, n, x0);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(err)); return 1; }
    std::vector<float> r(n);
    cudaMemcpy(r.data(), o, n * 4, cudaMemcpyDeviceToHost);
    // hypothesis: element (w, comp, q) of the box at byte offset o = ((w*4 + comp)*8 + q)*4, chunk XOR (w & 7)
    long bad = 0;
    for (int w = 0; w < bw; ++w) for (int c = 0; c < 4; ++c) for (int q = 0; q < 8; ++q) {
        int gw = x0 + w;
        float want = (gw < 0 || gw >= W || q >= Q) ? 0.f : h[(((size_t)1 * W + gw) * 4 + c) * Q + q];
        int chunk = (c * 2 + q / 4) ^ (w & 7);
        float got = r[w * 32 + chunk * 4 + (q & 3)];
        if (got != want) ++bad;
    }
    printf("Q=%d W=%d H=%d box w=%d x0=%d: hypothesis dense+xor(w&7): %ld mismatches of %d\n", Q, W, H, bw, x0, bad, n);
    for (int s = 0; s < 48 && s * 4 < n; ++s) {
        float v = r[s * 4];
        if (v < 0) { printf("chunk %2d: untouched\n", s); continue; }
        long g = (long)v;  // flat index in x[H][W][4][Q] of row h=1
        long rem = g - (long)1 * W * 4 * Q;
        printf("chunk %2d (line %d, slot %d): w=%ld comp=%ld q=%ld%s\n", s, s / 8, s % 8, rem / (4 * Q), (rem / Q) % 4, rem % Q,
               v == 0.f ? " (or zero fill)" : "");
    }
    return 0;
}
