/* qoracle_c.c -- TEST INFRASTRUCTURE, not product code: a plain-C restatement of the reference's quaternion conv / dense
 * forward, independent of the NumPy oracle (oracle/qoracle.py).  Only tests/ and __graft_entry__.build() may use it.
 *
 * It follows the reference LITERALLY in two steps, as its graph does:
 *   1. expand the stored kernel into the 4in_q x 4F real weight by slicing r,i,j,k, negating and concatenating
 *      (complexnn/conv.py:294-331 for the convolution, complexnn/dense.py:131-143 for the dense layer -- the dense
 *      block matrix is the transpose of the convolution's);
 *   2. run a real cross-correlation (no kernel flip) with TensorFlow's padding arithmetic (what K.conv1d / K.conv2d
 *      resolve to, complexnn/conv.py:309-334) or a real matmul (K.dot, complexnn/dense.py:149), add the bias on the
 *      channel axis (conv.py:336-341, dense.py:159-160) and apply relu when asked (conv.py:342-343).
 * fp32 in / out, fp64 accumulation.  Scalar loops, no dependencies:  gcc -O2 -shared -fPIC -o libqoracle_c.so qoracle_c.c
 */
#include <stdlib.h>
#include <string.h>

/* W_full[tap][a*in_q + q][b*F + f] from the stored kernel [tap][q][c*F + f]  (conv.py:327-331).
 * Row block a = input component, column block b = output component:
 *   b = r: [ f_r, -f_i, -f_j, -f_k ]   b = i: [ f_i, f_r, -f_k, f_j ]   b = j: [ f_j, f_k, f_r, -f_i ]   b = k: [ f_k, -f_j, f_i, f_r ]
 * (each list runs over a = r, i, j, k). */
static const int kIdx[4][4] = {{0, 1, 2, 3}, {1, 0, 3, 2}, {2, 3, 0, 1}, {3, 2, 1, 0}};      /* [b][a] -> sub-filter */
static const int kSgn[4][4] = {{1, -1, -1, -1}, {1, 1, -1, 1}, {1, 1, 1, -1}, {1, -1, 1, 1}}; /* [b][a] -> sign       */

static double* expand_conv(const float* kernel, int taps, int in_q, int F) {
    const int Ci = 4 * in_q, Co = 4 * F;
    double* w = (double*)malloc(sizeof(double) * (size_t)taps * Ci * Co);
    if (!w) return NULL;
    for (int t = 0; t < taps; ++t)
        for (int b = 0; b < 4; ++b)
            for (int a = 0; a < 4; ++a)
                for (int q = 0; q < in_q; ++q)
                    for (int f = 0; f < F; ++f)
                        w[((size_t)t * Ci + a * in_q + q) * Co + b * F + f] =
                            kSgn[b][a] * (double)kernel[((size_t)t * in_q + q) * Co + kIdx[b][a] * F + f];
    return w;
}

/* dense: rows fed by x_a are concat_out[...] of dense.py:139-142, i.e. W_dense[a][b] = W_conv[b][a] (transpose) */
static double* expand_dense(const float* kernel, int in_q, int Q) {
    const int Ci = 4 * in_q, Co = 4 * Q;
    double* w = (double*)malloc(sizeof(double) * (size_t)Ci * Co);
    if (!w) return NULL;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
            for (int q = 0; q < in_q; ++q)
                for (int f = 0; f < Q; ++f)
                    w[((size_t)a * in_q + q) * Co + b * Q + f] = kSgn[a][b] * (double)kernel[(size_t)q * Co + kIdx[a][b] * Q + f];
    return w;
}

/* tf.nn.convolution padding: VALID (0), SAME (1: total = max((ceil(n/s)-1)*s + (k-1)*d + 1 - n, 0), the odd one after),
 * causal (2: left pad d*(k-1), rank 1 only) */
static void resolve_axis(int n, int k, int s, int d, int padding, int* lo, int* out) {
    const int eff = (k - 1) * d + 1;
    if (padding == 0) {
        *lo = 0;
        *out = n >= eff ? (n - eff) / s + 1 : 0;
    } else if (padding == 1) {
        *out = (n + s - 1) / s;
        int total = (*out - 1) * s + eff - n;
        if (total < 0) total = 0;
        *lo = total / 2;
    } else {
        *out = (n + s - 1) / s;
        *lo = d * (k - 1);
    }
}

/* QuaternionConv1D forward, channels_last: x[B][L][4in_q], kernel[k][in_q][4F], bias[4F] or NULL -> y[B][Lo][4F].
 * Returns Lo (the caller sizes y with qoc_conv1d_out_len), or -1 on allocation failure. */
int qoc_conv1d_out_len(int L, int k, int stride, int dilation, int padding) {
    int lo, out;
    resolve_axis(L, k, stride, dilation, padding, &lo, &out);
    return out;
}

int qoc_conv1d_forward(const float* x, const float* kernel, const float* bias, float* y, int B, int L, int in_q, int F,
                       int k, int stride, int dilation, int padding, int relu) {
    int lo, Lo;
    resolve_axis(L, k, stride, dilation, padding, &lo, &Lo);
    const int Ci = 4 * in_q, Co = 4 * F;
    double* w = expand_conv(kernel, k, in_q, F);
    double* acc = (double*)malloc(sizeof(double) * Co);
    if (!w || !acc) {
        free(w);
        free(acc);
        return -1;
    }
    for (int n = 0; n < B; ++n)
        for (int o = 0; o < Lo; ++o) {
            for (int c = 0; c < Co; ++c) acc[c] = bias ? (double)bias[c] : 0.0;
            for (int t = 0; t < k; ++t) {
                const int p = o * stride - lo + t * dilation;
                if (p < 0 || p >= L) continue;
                const float* xr = x + ((size_t)n * L + p) * Ci;
                for (int ci = 0; ci < Ci; ++ci) {
                    const double xv = xr[ci];
                    const double* wr = w + ((size_t)t * Ci + ci) * Co;
                    for (int c = 0; c < Co; ++c) acc[c] += xv * wr[c];
                }
            }
            float* yr = y + ((size_t)n * Lo + o) * Co;
            for (int c = 0; c < Co; ++c) yr[c] = (float)((relu && acc[c] < 0.0) ? 0.0 : acc[c]);
        }
    free(w);
    free(acc);
    return Lo;
}

/* QuaternionConv2D forward, channels_first: x[B][4in_q][H][W], kernel[kh][kw][in_q][4F] -> y[B][4F][Ho][Wo]. */
void qoc_conv2d_out_shape(int H, int W, int kh, int kw, int sh, int sw, int dh, int dw, int padding, int* Ho, int* Wo) {
    int lo;
    resolve_axis(H, kh, sh, dh, padding, &lo, Ho);
    resolve_axis(W, kw, sw, dw, padding, &lo, Wo);
}

int qoc_conv2d_cf_forward(const float* x, const float* kernel, const float* bias, float* y, int B, int H, int W, int in_q,
                          int F, int kh, int kw, int sh, int sw, int dh, int dw, int padding, int relu) {
    int lo_h, lo_w, Ho, Wo;
    resolve_axis(H, kh, sh, dh, padding, &lo_h, &Ho);
    resolve_axis(W, kw, sw, dw, padding, &lo_w, &Wo);
    const int Ci = 4 * in_q, Co = 4 * F;
    double* w = expand_conv(kernel, kh * kw, in_q, F);
    if (!w) return -1;
    for (int n = 0; n < B; ++n)
        for (int c = 0; c < Co; ++c)
            for (int oh = 0; oh < Ho; ++oh)
                for (int ow = 0; ow < Wo; ++ow) {
                    double acc = bias ? (double)bias[c] : 0.0;
                    for (int i = 0; i < kh; ++i) {
                        const int ih = oh * sh - lo_h + i * dh;
                        if (ih < 0 || ih >= H) continue;
                        for (int j = 0; j < kw; ++j) {
                            const int iw = ow * sw - lo_w + j * dw;
                            if (iw < 0 || iw >= W) continue;
                            const double* wt = w + (size_t)(i * kw + j) * Ci * Co + c;
                            const float* xp = x + ((size_t)n * Ci * H + ih) * W + iw;
                            for (int ci = 0; ci < Ci; ++ci) acc += (double)xp[(size_t)ci * H * W] * wt[(size_t)ci * Co];
                        }
                    }
                    y[(((size_t)n * Co + c) * Ho + oh) * Wo + ow] = (float)((relu && acc < 0.0) ? 0.0 : acc);
                }
    free(w);
    return 0;
}

/* QuaternionDense forward: x[rows][4in_q], kernel[in_q][4Q] (units = 4Q), bias[4Q] or NULL -> y[rows][4Q]. */
int qoc_dense_forward(const float* x, const float* kernel, const float* bias, float* y, int rows, int in_q, int Q, int relu) {
    const int Ci = 4 * in_q, Co = 4 * Q;
    double* w = expand_dense(kernel, in_q, Q);
    double* acc = (double*)malloc(sizeof(double) * Co);
    if (!w || !acc) {
        free(w);
        free(acc);
        return -1;
    }
    for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < Co; ++c) acc[c] = bias ? (double)bias[c] : 0.0;
        for (int ci = 0; ci < Ci; ++ci) {
            const double xv = x[(size_t)r * Ci + ci];
            const double* wr = w + (size_t)ci * Co;
            for (int c = 0; c < Co; ++c) acc[c] += xv * wr[c];
        }
        for (int c = 0; c < Co; ++c) y[(size_t)r * Co + c] = (float)((relu && acc[c] < 0.0) ? 0.0 : acc[c]);
    }
    free(w);
    free(acc);
    return 0;
}
