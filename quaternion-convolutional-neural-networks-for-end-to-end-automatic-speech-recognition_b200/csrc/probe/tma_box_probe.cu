// Probe: which 4-D TMA boxes (no swizzle) load correctly on sm_100a.  usage: tma_box_probe W H Q C bw bh bq bc x0 y0
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../qnn_ptx.cuh"
#include "../qnn_tmap.h"
using namespace qnn;
using namespace qnn::ptx;

__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int n, int x0, int y0) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, n * 4);
        tma_load_4d(sm, &tm, &bar, x0, y0, 0, 0);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}

int main(int argc, char** argv) {
    if (argc < 11) return 2;
    int W = atoi(argv[1]), H = atoi(argv[2]), Q = atoi(argv[3]), C = atoi(argv[4]);
    uint32_t box[4] = {(uint32_t)atoi(argv[5]), (uint32_t)atoi(argv[6]), (uint32_t)atoi(argv[7]), (uint32_t)atoi(argv[8])};
    int x0 = atoi(argv[9]), y0 = atoi(argv[10]);
    size_t total = (size_t)W * H * Q * C;
    std::vector<float> h(total);
    for (size_t i = 0; i < total; ++i) h[i] = (float)(i + 1);
    float *d, *o;
    cudaMalloc(&d, total * 4);
    cudaMemcpy(d, h.data(), total * 4, cudaMemcpyHostToDevice);
    int n = box[0] * box[1] * box[2] * box[3];
    cudaMalloc(&o, n * 4);
    CUtensorMap tm;
    uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)Q, (uint64_t)C};
    uint64_t str[3] = {(uint64_t)W * 4, (uint64_t)W * H * 4, (uint64_t)W * H * Q * 4};
    int e = make_tmap_f32(&tm, d, 4, dims, str, box, false);
    if (e) { printf("encode failed %d\n", e); return 1; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<<<1, 128, 200 * 1024>>>(tm, o, n, x0, y0);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("box %u %u %u %u: %s\n", box[0], box[1], box[2], box[3], cudaGetErrorString(err)); return 1; }
    std::vector<float> r(n);
    cudaMemcpy(r.data(), o, n * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int c = 0; c < (int)box[3]; ++c) for (int q = 0; q < (int)box[2]; ++q) for (int y = 0; y < (int)box[1]; ++y) for (int x = 0; x < (int)box[0]; ++x) {
        int gx = x0 + x, gy = y0 + y;
        float want = (gx < 0 || gx >= W || gy < 0 || gy >= H || q >= Q || c >= C) ? 0.f : h[(((size_t)c * Q + q) * H + gy) * W + gx];
        if (r[((c * box[2] + q) * box[1] + y) * box[0] + x] != want) ++bad;
    }
    printf("box %u %u %u %u at (%d,%d): ok, %ld mismatches of %d\n", box[0], box[1], box[2], box[3], x0, y0, bad, n);
    return 0;
}
