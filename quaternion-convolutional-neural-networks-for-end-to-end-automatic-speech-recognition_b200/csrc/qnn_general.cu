// General quaternion convolution kernels (CUDA cores, fp32 FMA).  They implement the whole semantic surface of
// QuaternionConv / QuaternionDense -- rank 1..3, any stride / dilation / padding, both data formats, forward and the
// three gradients -- without ever forming the 4C_in x 4C_out expanded weight: each thread carries one quaternion
// accumulator and applies the Hamilton product directly on the four stored sub-filters.
// They are the fallback for shapes the tensor-core kernel (qnn_hamilton_tc.cu) does not take.
#include <algorithm>
#include "qnn_common.h"

namespace qnn {
namespace {

struct Q4 {
    float r, i, j, k;
};

// acc += w (x) x   (Hamilton product, weight on the left: complexnn/conv.py:327-331)
__device__ __forceinline__ void ham_acc(Q4& y, const Q4& x, const Q4& w) {
    y.r += x.r * w.r - x.i * w.i - x.j * w.j - x.k * w.k;
    y.i += x.r * w.i + x.i * w.r - x.j * w.k + x.k * w.j;
    y.j += x.r * w.j + x.i * w.k + x.j * w.r - x.k * w.i;
    y.k += x.r * w.k - x.i * w.j + x.j * w.i + x.k * w.r;
}
// acc += sum over the blocks that use sub-filter c of x_a * g_b with the table sign, conv convention
__device__ __forceinline__ void ham_wgrad(Q4& dw, const Q4& x, const Q4& g) {
    dw.r += x.r * g.r + x.i * g.i + x.j * g.j + x.k * g.k;
    dw.i += x.r * g.i - x.i * g.r + x.j * g.k - x.k * g.j;
    dw.j += x.r * g.j - x.i * g.k - x.j * g.r + x.k * g.i;
    dw.k += x.r * g.k + x.i * g.j - x.j * g.i - x.k * g.r;
}

struct Idx {
    // element offset of (n, spatial position pos, real channel c) for a tensor with C channels and S positions
    int64_t S, C;
    int cf;
    __device__ __forceinline__ int64_t operator()(int64_t n, int64_t pos, int64_t c) const {
        return cf ? (n * C + c) * S + pos : (n * S + pos) * C + c;
    }
};

__device__ __forceinline__ float dact(float dy, float y, int act) { return act == QNN_ACT_RELU ? (y > 0.f ? dy : 0.f) : dy; }

// ---------------------------------------------------------------------------------------------------------------------
// forward: one thread = one output quaternion (n, p, f)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_general_fwd(Geom g, const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ y) {
    const int64_t P = (int64_t)g.out_sp[0] * g.out_sp[1] * g.out_sp[2];
    const int64_t S = (int64_t)g.in_sp[0] * g.in_sp[1] * g.in_sp[2];
    const int64_t total = (int64_t)g.batch * P * g.F;
    const int F = g.F, Q = g.in_q;
    const Idx xi{S, 4 * (int64_t)Q, g.channels_first}, yi{P, 4 * (int64_t)F, g.channels_first};
    const float cs = g.conj_w ? -1.f : 1.f;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t n, p;
        int f;
        if (g.channels_first) {
            p = t % P;
            f = (int)((t / P) % F);
            n = t / (P * F);
        } else {
            f = (int)(t % F);
            p = (t / F) % P;
            n = t / (P * F);
        }
        const int ow = (int)(p % g.out_sp[2]), oh = (int)((p / g.out_sp[2]) % g.out_sp[1]),
                  od = (int)(p / ((int64_t)g.out_sp[2] * g.out_sp[1]));
        Q4 acc{0.f, 0.f, 0.f, 0.f};
        for (int kd = 0; kd < g.k[0]; ++kd) {
            const int id = od * g.s[0] - g.pad_lo[0] + kd * g.d[0];
            if (id < 0 || id >= g.in_sp[0]) continue;
            for (int kh = 0; kh < g.k[1]; ++kh) {
                const int ih = oh * g.s[1] - g.pad_lo[1] + kh * g.d[1];
                if (ih < 0 || ih >= g.in_sp[1]) continue;
                for (int kw = 0; kw < g.k[2]; ++kw) {
                    const int iw = ow * g.s[2] - g.pad_lo[2] + kw * g.d[2];
                    if (iw < 0 || iw >= g.in_sp[2]) continue;
                    const int64_t ipos = ((int64_t)id * g.in_sp[1] + ih) * g.in_sp[2] + iw;
                    const int tap = (kd * g.k[1] + kh) * g.k[2] + kw;
                    const float* wp = w + (int64_t)tap * Q * 4 * F + f;
                    for (int q = 0; q < Q; ++q) {
                        const Q4 xv{__ldg(x + xi(n, ipos, q)), __ldg(x + xi(n, ipos, Q + q)),
                                    __ldg(x + xi(n, ipos, 2 * Q + q)), __ldg(x + xi(n, ipos, 3 * Q + q))};
                        const float* wq = wp + (int64_t)q * 4 * F;
                        const Q4 wv{__ldg(wq), cs * __ldg(wq + F), cs * __ldg(wq + 2 * F), cs * __ldg(wq + 3 * F)};
                        ham_acc(acc, xv, wv);
                    }
                }
            }
        }
        if (bias) {
            acc.r += __ldg(bias + f);
            acc.i += __ldg(bias + F + f);
            acc.j += __ldg(bias + 2 * F + f);
            acc.k += __ldg(bias + 3 * F + f);
        }
        y[yi(n, p, f)] = act_apply(acc.r, g.act);
        y[yi(n, p, F + f)] = act_apply(acc.i, g.act);
        y[yi(n, p, 2 * F + f)] = act_apply(acc.j, g.act);
        y[yi(n, p, 3 * F + f)] = act_apply(acc.k, g.act);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// dgrad: one thread = one input quaternion (n, ipos, q).  dx = conj-table Hamilton product of dz with the taps that
// reach this input position (conv-dgrad uses the dense table and vice versa).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_general_dgrad(Geom g, const float* __restrict__ w, const float* __restrict__ y,
                                                       const float* __restrict__ dy, float* __restrict__ dx) {
    const int64_t P = (int64_t)g.out_sp[0] * g.out_sp[1] * g.out_sp[2];
    const int64_t S = (int64_t)g.in_sp[0] * g.in_sp[1] * g.in_sp[2];
    const int F = g.F, Q = g.in_q;
    const int64_t total = (int64_t)g.batch * S * Q;
    const Idx xi{S, 4 * (int64_t)Q, g.channels_first}, yi{P, 4 * (int64_t)F, g.channels_first};
    const float cs = g.conj_w ? 1.f : -1.f;  // opposite convention of the forward
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t n, ip;
        int q;
        if (g.channels_first) {
            ip = t % S;
            q = (int)((t / S) % Q);
            n = t / (S * Q);
        } else {
            q = (int)(t % Q);
            ip = (t / Q) % S;
            n = t / (S * Q);
        }
        const int iw = (int)(ip % g.in_sp[2]), ih = (int)((ip / g.in_sp[2]) % g.in_sp[1]),
                  id = (int)(ip / ((int64_t)g.in_sp[2] * g.in_sp[1]));
        Q4 acc{0.f, 0.f, 0.f, 0.f};
        for (int kd = 0; kd < g.k[0]; ++kd) {
            int nd = id + g.pad_lo[0] - kd * g.d[0];
            if (nd < 0 || nd % g.s[0]) continue;
            nd /= g.s[0];
            if (nd >= g.out_sp[0]) continue;
            for (int kh = 0; kh < g.k[1]; ++kh) {
                int nh = ih + g.pad_lo[1] - kh * g.d[1];
                if (nh < 0 || nh % g.s[1]) continue;
                nh /= g.s[1];
                if (nh >= g.out_sp[1]) continue;
                for (int kw = 0; kw < g.k[2]; ++kw) {
                    int nw = iw + g.pad_lo[2] - kw * g.d[2];
                    if (nw < 0 || nw % g.s[2]) continue;
                    nw /= g.s[2];
                    if (nw >= g.out_sp[2]) continue;
                    const int64_t op = ((int64_t)nd * g.out_sp[1] + nh) * g.out_sp[2] + nw;
                    const int tap = (kd * g.k[1] + kh) * g.k[2] + kw;
                    const float* wq = w + ((int64_t)tap * Q + q) * 4 * F;
                    for (int f = 0; f < F; ++f) {
                        const int64_t o0 = yi(n, op, f), o1 = yi(n, op, F + f), o2 = yi(n, op, 2 * F + f),
                                      o3 = yi(n, op, 3 * F + f);
                        const Q4 gz{dact(__ldg(dy + o0), __ldg(y + o0), g.act), dact(__ldg(dy + o1), __ldg(y + o1), g.act),
                                    dact(__ldg(dy + o2), __ldg(y + o2), g.act), dact(__ldg(dy + o3), __ldg(y + o3), g.act)};
                        const Q4 wv{__ldg(wq + f), cs * __ldg(wq + F + f), cs * __ldg(wq + 2 * F + f),
                                    cs * __ldg(wq + 3 * F + f)};
                        ham_acc(acc, gz, wv);
                    }
                }
            }
        }
        dx[xi(n, ip, q)] = acc.r;
        dx[xi(n, ip, Q + q)] = acc.i;
        dx[xi(n, ip, 2 * Q + q)] = acc.j;
        dx[xi(n, ip, 3 * Q + q)] = acc.k;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// wgrad: thread = one stored weight quaternion (tap, q, f); blockIdx.y splits the (n, out position) reduction; the 16
// blocks of dL/dW_full are folded into the 4 stored sub-filters on the fly and added with atomics.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_general_wgrad(Geom g, const float* __restrict__ x, const float* __restrict__ y,
                                                       const float* __restrict__ dy, float* __restrict__ dw) {
    const int64_t P = (int64_t)g.out_sp[0] * g.out_sp[1] * g.out_sp[2];
    const int64_t S = (int64_t)g.in_sp[0] * g.in_sp[1] * g.in_sp[2];
    const int F = g.F, Q = g.in_q;
    const int taps = g.k[0] * g.k[1] * g.k[2];
    const int64_t nw = (int64_t)taps * Q * F;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nw) return;
    const int f = (int)(t % F), q = (int)((t / F) % Q), tap = (int)(t / ((int64_t)F * Q));
    const int kw = tap % g.k[2], kh = (tap / g.k[2]) % g.k[1], kd = tap / (g.k[2] * g.k[1]);
    const Idx xi{S, 4 * (int64_t)Q, g.channels_first}, yi{P, 4 * (int64_t)F, g.channels_first};
    const int64_t rows = (int64_t)g.batch * P;
    const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    Q4 acc{0.f, 0.f, 0.f, 0.f};
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t n = r / P, op = r % P;
        const int ow = (int)(op % g.out_sp[2]), oh = (int)((op / g.out_sp[2]) % g.out_sp[1]),
                  od = (int)(op / ((int64_t)g.out_sp[2] * g.out_sp[1]));
        const int id = od * g.s[0] - g.pad_lo[0] + kd * g.d[0], ih = oh * g.s[1] - g.pad_lo[1] + kh * g.d[1],
                  iw = ow * g.s[2] - g.pad_lo[2] + kw * g.d[2];
        if (id < 0 || id >= g.in_sp[0] || ih < 0 || ih >= g.in_sp[1] || iw < 0 || iw >= g.in_sp[2]) continue;
        const int64_t ip = ((int64_t)id * g.in_sp[1] + ih) * g.in_sp[2] + iw;
        const Q4 xv{__ldg(x + xi(n, ip, q)), __ldg(x + xi(n, ip, Q + q)), __ldg(x + xi(n, ip, 2 * Q + q)),
                    __ldg(x + xi(n, ip, 3 * Q + q))};
        const int64_t o0 = yi(n, op, f), o1 = yi(n, op, F + f), o2 = yi(n, op, 2 * F + f), o3 = yi(n, op, 3 * F + f);
        const Q4 gz{dact(__ldg(dy + o0), __ldg(y + o0), g.act), dact(__ldg(dy + o1), __ldg(y + o1), g.act),
                    dact(__ldg(dy + o2), __ldg(y + o2), g.act), dact(__ldg(dy + o3), __ldg(y + o3), g.act)};
        ham_wgrad(acc, xv, gz);
    }
    const float cs = g.conj_w ? -1.f : 1.f;
    float* o = dw + ((int64_t)tap * Q + q) * 4 * F + f;
    atomicAdd(o, acc.r);
    atomicAdd(o + F, cs * acc.i);
    atomicAdd(o + 2 * F, cs * acc.j);
    atomicAdd(o + 3 * F, cs * acc.k);
}

// dbias[c] = sum over (n, position) of dz[., c]
__global__ void __launch_bounds__(256) k_general_bgrad(Geom g, const float* __restrict__ y, const float* __restrict__ dy,
                                                       float* __restrict__ db) {
    const int64_t P = (int64_t)g.out_sp[0] * g.out_sp[1] * g.out_sp[2];
    const int C = 4 * g.F;
    const Idx yi{P, C, g.channels_first};
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int64_t rows = (int64_t)g.batch * P;
    const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    float acc = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t o = yi(r / P, r % P, c);
        acc += dact(__ldg(dy + o), __ldg(y + o), g.act);
    }
    atomicAdd(db + c, acc);
}

// ---------------------------------------------------------------------------------------------------------------------
// helpers of the tensor-core backward path (channels_last rows x C, C = 4F)
// ---------------------------------------------------------------------------------------------------------------------
// dz = dy * act'(y) (relu: y > 0) and dbias[c] += sum over rows of dz[., c] in ONE pass over (y, dy).
// blockDim = (bx, by): bx float4-columns (looped when C/4 > bx), by row lanes; grid.x = row chunks.
__global__ void __launch_bounds__(256) k_dz_bgrad(const float4* __restrict__ y, const float4* __restrict__ dy,
                                                  float4* __restrict__ dz, float* __restrict__ db, long long rows, int C4,
                                                  int relu) {
    extern __shared__ float4 red[];  // [by][bx]
    const long long per = (rows + gridDim.x - 1) / gridDim.x;
    const long long r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
    // uniform trip count (the loop body holds __syncthreads): threads past the last column run masked iterations
    const int c4_end = ((C4 + (int)blockDim.x - 1) / (int)blockDim.x) * (int)blockDim.x;
    for (int c4i = threadIdx.x; c4i < c4_end; c4i += blockDim.x) {
        const bool live = c4i < C4;
        const int c4 = live ? c4i : C4 - 1;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // four rows per iteration: eight independent 16-byte loads in flight per thread
        long long r = r0 + threadIdx.y;
        const long long step = blockDim.y;
        const long long r_end = live ? r1 : r0;  // masked threads skip the row loops
        for (; r + 3 * step < r_end; r += 4 * step) {
            float4 g[4], v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) g[u] = __ldg(dy + (r + u * step) * C4 + c4);
            if (relu) {
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldg(y + (r + u * step) * C4 + c4);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    g[u].x = v[u].x > 0.f ? g[u].x : 0.f;
                    g[u].y = v[u].y > 0.f ? g[u].y : 0.f;
                    g[u].z = v[u].z > 0.f ? g[u].z : 0.f;
                    g[u].w = v[u].w > 0.f ? g[u].w : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (dz) dz[(r + u * step) * C4 + c4] = g[u];
                acc.x += g[u].x;
                acc.y += g[u].y;
                acc.z += g[u].z;
                acc.w += g[u].w;
            }
        }
        for (; r < r_end; r += step) {
            const long long o = r * C4 + c4;
            float4 g = __ldg(dy + o);
            if (relu) {
                const float4 v = __ldg(y + o);
                g.x = v.x > 0.f ? g.x : 0.f;
                g.y = v.y > 0.f ? g.y : 0.f;
                g.z = v.z > 0.f ? g.z : 0.f;
                g.w = v.w > 0.f ? g.w : 0.f;
            }
            if (dz) dz[o] = g;
            acc.x += g.x;
            acc.y += g.y;
            acc.z += g.z;
            acc.w += g.w;
        }
        if (db) {
            red[threadIdx.y * blockDim.x + threadIdx.x] = acc;
            __syncthreads();
            if (threadIdx.y == 0 && live) {
                for (int j = 1; j < (int)blockDim.y; ++j) {
                    const float4 o = red[j * blockDim.x + threadIdx.x];
                    acc.x += o.x;
                    acc.y += o.y;
                    acc.z += o.z;
                    acc.w += o.w;
                }
                atomicAdd(db + 4 * c4 + 0, acc.x);
                atomicAdd(db + 4 * c4 + 1, acc.y);
                atomicAdd(db + 4 * c4 + 2, acc.z);
                atomicAdd(db + 4 * c4 + 3, acc.w);
            }
            __syncthreads();
        }
    }
}

// channels_first twin: tensors are [n][C][S] (S contiguous positions per channel).  One block walks whole (n, c) rows:
// dz = dy * act'(y) with 16-byte accesses where the row is aligned, block-reduced row sums added to dbias[c].
__global__ void __launch_bounds__(256) k_dz_bgrad_cf(const float* __restrict__ y, const float* __restrict__ dy,
                                                     float* __restrict__ dz, float* __restrict__ db, long long rows, int C,
                                                     long long S, int relu) {
    __shared__ float red[8];
    const bool vec = (S % 4) == 0;  // every row then starts on a 16-byte boundary (the base pointers are 16-byte aligned)
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const long long base = row * S;
        float acc = 0.f;
        if (vec) {
            const float4* y4 = reinterpret_cast<const float4*>(y + base);
            const float4* g4 = reinterpret_cast<const float4*>(dy + base);
            float4* z4 = dz ? reinterpret_cast<float4*>(dz + base) : nullptr;
            for (long long i = threadIdx.x; i < S / 4; i += blockDim.x) {
                float4 g = __ldg(g4 + i);
                if (relu) {
                    const float4 v = __ldg(y4 + i);
                    g.x = v.x > 0.f ? g.x : 0.f;
                    g.y = v.y > 0.f ? g.y : 0.f;
                    g.z = v.z > 0.f ? g.z : 0.f;
                    g.w = v.w > 0.f ? g.w : 0.f;
                }
                if (z4) z4[i] = g;
                acc += (g.x + g.y) + (g.z + g.w);
            }
        } else {
            for (long long i = threadIdx.x; i < S; i += blockDim.x) {
                float g = __ldg(dy + base + i);
                if (relu && !(__ldg(y + base + i) > 0.f)) g = 0.f;
                if (dz) dz[base + i] = g;
                acc += g;
            }
        }
        if (db) {  // block-uniform
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
            __syncthreads();
            if (threadIdx.x == 0) {
                float t = 0.f;
                for (int j = 0; j < (int)(blockDim.x >> 5); ++j) t += red[j];
                atomicAdd(db + (int)(row % C), t);
            }
            __syncthreads();
        }
    }
}

// [n][C][S] (channels_first) -> [n][S][C] (channels_last), 32 x 32 tiles through shared memory: both sides coalesced
__global__ void __launch_bounds__(256) k_cf_to_cl(const float* __restrict__ in, float* __restrict__ out, int C, long long S,
                                                  long long tiles_s, long long n_tiles) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const long long tiles_c = (C + 31) / 32;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long ts = t % tiles_s, tc = (t / tiles_s) % tiles_c, n = t / (tiles_s * tiles_c);
        const float* src = in + (size_t)n * C * S;
        float* dst = out + (size_t)n * C * S;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const long long c = tc * 32 + ty + j, sidx = ts * 32 + tx;
            if (c < C && sidx < S) tile[ty + j][tx] = __ldg(src + c * S + sidx);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const long long sidx = ts * 32 + ty + j, c = tc * 32 + tx;
            if (c < C && sidx < S) dst[sidx * C + c] = tile[tx][ty + j];
        }
        __syncthreads();
    }
}

inline int grid_for(int64_t total, int block) {
    int64_t b = (total + block - 1) / block;
    const int64_t cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int general_forward(const Geom& g, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
    const int64_t P = (int64_t)g.out_sp[0] * g.out_sp[1] * g.out_sp[2];
    const int64_t total = (int64_t)g.batch * P * g.F;
    if (total == 0) return QNN_OK;
    k_general_fwd<<<grid_for(total, 256), 256, 0, st>>>(g, x, w, bias, y);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("general forward launch failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    return QNN_OK;
}

int dz_bgrad(const float* y, const float* dy, float* dz, float* db, long long rows, int C, int relu, cudaStream_t st) {
    if (C % 4) {
        set_error("dz_bgrad needs a channel count that is a multiple of 4");
        return QNN_E_INVALID;
    }
    cudaError_t e;
    if (db && (e = cudaMemsetAsync(db, 0, (size_t)C * sizeof(float), st)) != cudaSuccess) {
        set_error("dbias memset failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    if (rows == 0) return QNN_OK;
    const int C4 = C / 4;
    int bx = 1;
    while (bx * 2 <= C4 && bx * 2 <= 256) bx *= 2;  // power of two <= min(C/4, 256); wider rows loop over columns
    if (C4 <= 256 && 256 % C4 == 0) bx = C4;
    const int by = std::max(1, 256 / bx);
    const int grid = (int)std::min<long long>((rows + by - 1) / by, 148 * 8);
    k_dz_bgrad<<<grid, dim3(bx, by), (size_t)bx * by * sizeof(float4), st>>>(
        reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dz), db, rows, C4,
        relu);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("dz/bgrad launch failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    return QNN_OK;
}

int dz_bgrad_cf(const float* y, const float* dy, float* dz, float* db, int batch, int C, long long S, int relu,
                cudaStream_t st) {
    cudaError_t e;
    if (db && (e = cudaMemsetAsync(db, 0, (size_t)C * sizeof(float), st)) != cudaSuccess) {
        set_error("dbias memset failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    const long long rows = (long long)batch * C;
    if (rows == 0 || S == 0) return QNN_OK;
    k_dz_bgrad_cf<<<(unsigned)std::min<long long>(rows, 148LL * 8), 256, 0, st>>>(y, dy, dz, db, rows, C, S, relu);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("dz/bgrad (channels_first) launch failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    return QNN_OK;
}

int cf_to_cl(const float* in, float* out, int batch, int C, long long S, cudaStream_t st) {
    const long long tiles_s = (S + 31) / 32, n_tiles = (long long)batch * ((C + 31) / 32) * tiles_s;
    if (n_tiles == 0) return QNN_OK;
    k_cf_to_cl<<<(unsigned)std::min<long long>(n_tiles, 148LL * 16), 256, 0, st>>>(in, out, C, S, tiles_s, n_tiles);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("layout transposition launch failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    return QNN_OK;
}

int general_backward(const Geom& g, const float* x, const float* w, const float* y, const float* dy, float* dx, float* dw,
                     float* db, cudaStream_t st) {
    const int64_t P = (int64_t)g.out_sp[0] * g.out_sp[1] * g.out_sp[2];
    const int64_t S = (int64_t)g.in_sp[0] * g.in_sp[1] * g.in_sp[2];
    const int taps = g.k[0] * g.k[1] * g.k[2];
    const int64_t rows = (int64_t)g.batch * P;
    cudaError_t e;
    if (dx) {
        const int64_t total = (int64_t)g.batch * S * g.in_q;
        if (total) {
            k_general_dgrad<<<grid_for(total, 256), 256, 0, st>>>(g, w, y, dy, dx);
            count_launch();
        }
    }
    if (dw) {
        const int64_t nw = (int64_t)taps * g.in_q * g.F;
        if ((e = cudaMemsetAsync(dw, 0, nw * 4 * sizeof(float), st)) != cudaSuccess) goto fail;
        if (rows && nw) {
            const int gx = (int)((nw + 255) / 256);
            int64_t split = (148LL * 8 + gx - 1) / gx;
            if (split > rows) split = rows;
            if (split < 1) split = 1;
            if (split > 65535) split = 65535;
            k_general_wgrad<<<dim3(gx, (unsigned)split), 256, 0, st>>>(g, x, y, dy, dw);
            count_launch();
        }
    }
    if (db) {
        if ((e = cudaMemsetAsync(db, 0, 4 * g.F * sizeof(float), st)) != cudaSuccess) goto fail;
        if (rows) {
            const int gx = (4 * g.F + 255) / 256;
            int64_t split = (148LL * 4 + gx - 1) / gx;
            if (split > rows) split = rows;
            if (split > 65535) split = 65535;
            k_general_bgrad<<<dim3(gx, (unsigned)split), 256, 0, st>>>(g, y, dy, db);
            count_launch();
        }
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) return QNN_OK;
fail:
    set_error("general backward failed: %s", cudaGetErrorString(e));
    return QNN_E_CUDA;
}

}  // namespace qnn
