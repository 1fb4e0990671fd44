// Small-K quaternion convolution / dense forward on the CUDA cores (fp32 FMA): layers with fewer than four quaternion
// input channels -- the first layer of the reference's DECODA model (models/example_model.py:25: QuaternionConv1D(32, 3)
// on x[B, 250, 4], in_q = 1) -- where the contraction (taps * 4 * in_q = 12 reals) is far too short for a tensor-core
// tile and the layer is bound by writing y.
//
// channels_last rows, taps along the sequence, stride 1 ("the 1D / stride-1 im2col path").  One WARP owns a run of PC
// consecutive output positions of one sequence and 32 filters:
//   * the input window of the run (PC + halo positions x 4*in_q channels, contiguous in memory) is fetched with ONE
//     coalesced 16-byte load per lane (128 floats per warp, out-of-sequence positions read as zero = the padding);
//   * every tap of every position is served from those registers: the value (position + tap * dilation, component a,
//     channel q) lives in a known lane and register (compile-time offsets from the position's first lane: in_q is a
//     template parameter) and is broadcast with ONE warp shuffle -- the im2col row is never materialised, x is read from
//     memory once however many taps reuse it;
//   * lanes are FILTERS: each lane keeps the four accumulators (r, i, j, k) of its filter, reads its four sub-filter
//     weights of a (tap, q) from shared memory (consecutive lanes -> consecutive words, conflict-free; the block holds
//     the stored [tap][q][r|i|j|k][F] kernel as is) and applies the Hamilton product directly (16 FMA per quaternion
//     MAC) -- no expanded 4in_q x 4F weight (complexnn/conv.py:327-331);
//   * bias + activation fused; the four component rows of y are written as coalesced 128-byte segments.
#include <algorithm>
#include "qnn_common.h"

namespace qnn {
namespace {

constexpr int kWarps = 8;
constexpr int kWindow = 128;  // floats of x a warp holds in registers (one float4 per lane)

struct SmallK {
    int batch, L, Lo, in_q, F, taps, dil, pad_lo;
    int C;               // 4 * in_q real input channels
    int pc;              // output positions per run
    int runs_per_seq, n_fg;
    long long n_tasks;
    int act, conj_w, has_bias;
};

template <int M>
__device__ __forceinline__ float f4_get(const float4& v) {
    return M == 0 ? v.x : M == 1 ? v.y : M == 2 ? v.z : v.w;
}

// The 4*Q floats [r(Q) | i(Q) | j(Q) | k(Q)] of the position whose first float4 lives in lane `lane0`: float m sits in
// lane lane0 + m/4, component m%4 -- both compile-time offsets, so each value is ONE shuffle (no branch).
template <int Q, int M>
struct Gather {
    __device__ static __forceinline__ void run(const float4& v, int lane0, float (&xv)[4 * Q]) {
        xv[M] = __shfl_sync(0xffffffffu, f4_get<M & 3>(v), lane0 + (M >> 2));
        Gather<Q, M + 1>::run(v, lane0, xv);
    }
};
template <int Q>
struct Gather<Q, 4 * Q> {
    __device__ static __forceinline__ void run(const float4&, int, float (&)[4 * Q]) {}
};

// TAPS > 0: the lane's taps * Q * 4 sub-filter weights live in registers for the whole run (ncu on the first version:
// the kernel is issue-bound, 181 instructions per position, a third of them shared-memory weight loads and sign
// multiplies); TAPS == 0: any tap count, weights re-read from shared memory per position.
template <int Q, int TAPS>
__global__ void __launch_bounds__(kWarps * 32) k_smallk_fwd(const SmallK p, const float* __restrict__ x,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ y) {
    extern __shared__ float smem[];
    float* w_s = smem;                                   // [tap][q][c][F], as stored, i / j / k already carrying the
    float* b_s = smem + (size_t)p.taps * Q * 4 * p.F;    // dense-convention sign;  bias [c][F]
    const int n_w = p.taps * Q * 4 * p.F, F = p.F;
    // dense convention y = conj(W) (x) x (complexnn/dense.py:139-143): the imaginary sub-filters enter negated
    for (int i = threadIdx.x; i < n_w; i += blockDim.x) {
        const float v = __ldg(w + i);
        w_s[i] = (p.conj_w && ((i / F) & 3)) ? -v : v;
    }
    for (int i = threadIdx.x; i < 4 * F; i += blockDim.x) b_s[i] = p.has_bias ? __ldg(bias + i) : 0.f;
    __syncthreads();

    constexpr int C = 4 * Q;
    constexpr int NT = TAPS > 0 ? TAPS : 1;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5), n_warps = (long long)gridDim.x * kWarps;
    const bool relu = p.act == QNN_ACT_RELU, generic_act = p.act != QNN_ACT_RELU && p.act != QNN_ACT_LINEAR;
    // this lane's slice of the window: floats [4*lane, 4*lane + 4) = channels wch..wch+3 of position woff (C % 4 == 0)
    const int woff = (4 * lane) / C, wch = (4 * lane) % C;
    int fg_loaded = -1;
    float wreg[NT * Q][4];
    for (long long task = warp0; task < p.n_tasks; task += n_warps) {
        const int fg = (int)(task % p.n_fg);
        const long long t2 = task / p.n_fg;
        const int run = (int)(t2 % p.runs_per_seq);
        const long long n = t2 / p.runs_per_seq;
        const int t0 = run * p.pc;
        const int pos = t0 - p.pad_lo + woff;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pos >= 0 && pos < p.L) v = __ldg(reinterpret_cast<const float4*>(x + ((size_t)n * p.L + pos) * C + wch));
        const int f = fg * 32 + lane;
        const bool fv = f < F;
        const int fc = fv ? f : F - 1;
        if (TAPS > 0 && fg != fg_loaded) {  // (warp-uniform) this lane's filter changed: reload its weights
#pragma unroll
            for (int i = 0; i < NT * Q; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) wreg[i][c] = w_s[((size_t)i * 4 + c) * F + fc];
            fg_loaded = fg;
        }
        const float b0 = b_s[fc], b1 = b_s[F + fc], b2 = b_s[2 * F + fc], b3 = b_s[3 * F + fc];
        const int n_pos = min(p.pc, p.Lo - t0);
        float* o = y + ((size_t)n * p.Lo + t0) * 4 * F + f;
        for (int j = 0; j < n_pos; ++j, o += 4 * F) {
            float yr = b0, yi = b1, yj = b2, yk = b3;
            auto qmac = [&](const float (&xv)[C], int q, float wr, float wi, float wj, float wk) {
                const float xr = xv[q], xi = xv[Q + q], xj = xv[2 * Q + q], xk = xv[3 * Q + q];
                // y += w (x) x  (Hamilton product, weight on the left: complexnn/conv.py:327-331)
                yr += xr * wr - xi * wi - xj * wj - xk * wk;
                yi += xr * wi + xi * wr - xj * wk + xk * wj;
                yj += xr * wj + xi * wk + xj * wr - xk * wi;
                yk += xr * wk - xi * wj + xj * wi + xk * wr;
            };
            if (TAPS > 0) {
#pragma unroll
                for (int tap = 0; tap < NT; ++tap) {
                    float xv[C];
                    Gather<Q, 0>::run(v, (j + tap * p.dil) * Q, xv);  // the position's first float4 is in lane position * C / 4
#pragma unroll
                    for (int q = 0; q < Q; ++q)
                        qmac(xv, q, wreg[tap * Q + q][0], wreg[tap * Q + q][1], wreg[tap * Q + q][2], wreg[tap * Q + q][3]);
                }
            } else {
                for (int tap = 0; tap < p.taps; ++tap) {
                    float xv[C];
                    Gather<Q, 0>::run(v, (j + tap * p.dil) * Q, xv);
                    const float* wt = w_s + (size_t)tap * Q * 4 * F + fc;
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const float* wq = wt + (size_t)q * 4 * F;
                        qmac(xv, q, wq[0], wq[F], wq[2 * F], wq[3 * F]);
                    }
                }
            }
            if (relu) {
                yr = fmaxf(yr, 0.f), yi = fmaxf(yi, 0.f), yj = fmaxf(yj, 0.f), yk = fmaxf(yk, 0.f);
            } else if (generic_act) {
                yr = act_apply(yr, p.act), yi = act_apply(yi, p.act), yj = act_apply(yj, p.act), yk = act_apply(yk, p.act);
            }
            if (fv) o[0] = yr, o[F] = yi, o[2 * F] = yj, o[3 * F] = yk;
        }
    }
}

typedef void (*SmallKKernel)(const SmallK, const float*, const float*, const float*, float*);
template <int Q>
SmallKKernel pick_q(int taps) {
    switch (taps) {
        case 1: return k_smallk_fwd<Q, 1>;
        case 2: return k_smallk_fwd<Q, 2>;
        case 3: return k_smallk_fwd<Q, 3>;
        case 4: return k_smallk_fwd<Q, 4>;
        case 5: return k_smallk_fwd<Q, 5>;
        default: return k_smallk_fwd<Q, 0>;
    }
}
SmallKKernel pick_kernel(int in_q, int taps) { return in_q == 1 ? pick_q<1>(taps) : in_q == 2 ? pick_q<2>(taps) : pick_q<3>(taps); }

}  // namespace

SmallKPlan smallk_plan(const Geom& g, int rank) {
    SmallKPlan pl{};
    pl.ok = 0;
    auto no = [&](const char* why) {
        pl.why = why;
        return pl;
    };
    if (g.channels_first) return no("channels_first layout");
    if (rank != 1) return no("rank > 1");
    if (g.s[2] != 1) return no("stride != 1");
    if (g.in_q >= 4) return no("4 or more quaternion input channels (tensor-core territory)");
    if (g.out_sp[2] < 1 || g.batch < 1) return no("empty problem");
    const int C = 4 * g.in_q, halo = (g.k[2] - 1) * g.d[2];
    const int pc = kWindow / C - halo;
    if (pc < 1) return no("halo does not fit the 128-float register window");
    const size_t smem = ((size_t)g.k[2] * g.in_q * 4 * g.F + 4 * (size_t)g.F) * sizeof(float);
    if (smem > 200 * 1024) return no("stored kernel does not fit in shared memory");
    pl.ok = 1;
    pl.pc = pc;
    pl.smem_bytes = smem;
    pl.why = "";
    return pl;
}

int smallk_forward(const Geom& g, int rank, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
    const SmallKPlan pl = smallk_plan(g, rank);
    if (!pl.ok) {
        set_error("small-K kernel does not take this shape: %s", pl.why);
        return QNN_E_UNSUPPORTED;
    }
    if (reinterpret_cast<uintptr_t>(x) & 15) {
        set_error("small-K kernel needs a 16-byte aligned x");
        return QNN_E_UNSUPPORTED;
    }
    SmallK p{};
    p.batch = g.batch;
    p.L = g.in_sp[2];
    p.Lo = g.out_sp[2];
    p.in_q = g.in_q;
    p.F = g.F;
    p.taps = g.k[2];
    p.dil = g.d[2];
    p.pad_lo = g.pad_lo[2];
    p.C = 4 * g.in_q;
    p.pc = pl.pc;
    p.runs_per_seq = (p.Lo + p.pc - 1) / p.pc;
    p.n_fg = (g.F + 31) / 32;
    p.n_tasks = (long long)g.batch * p.runs_per_seq * p.n_fg;
    p.act = g.act;
    p.conj_w = g.conj_w;
    p.has_bias = bias != nullptr;
    SmallKKernel kern = pick_kernel(g.in_q, g.k[2]);
    if (pl.smem_bytes > 48 * 1024)
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)pl.smem_bytes)) return rc;
    // persistent warps: a multiple of the SM count, enough resident warps to cover the load latency, never more blocks
    // than tasks (every block re-reads the stored kernel into its shared memory)
    const long long want = (p.n_tasks + kWarps - 1) / kWarps;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, 4LL * num_sms()));
    kern<<<grid, kWarps * 32, pl.smem_bytes, st>>>(p, x, w, bias, y);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("small-K kernel launch failed: %s", cudaGetErrorString(e));
        return QNN_E_CUDA;
    }
    return QNN_OK;
}

}  // namespace qnn
