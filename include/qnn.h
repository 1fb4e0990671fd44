/* qnn.h -- C ABI of the B200-native quaternion convolution / dense hot path (libqnn_b200.so).
 *
 * The reference (Orkis-Research QCNN, /root/reference) has no FFI: the path sits behind the Keras 2 `Layer` protocol and
 * all arithmetic is delegated to Keras-backend calls.  Each entry point below replaces one such delegation; the
 * Python layer mirror (`complexnn` inside the *_b200 package) binds them with ctypes.
 *
 * Conventions: every function returns 0 (QNN_OK) or a negative qnn_status; nothing throws; caller owns every buffer;
 * pointers are DEVICE pointers unless the name says `_host`; work is enqueued asynchronously on `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream) and no device-pointer call synchronises the device
 * (the `_host` calls return after their result landed, i.e. they synchronise `stream` only);
 * qnn_last_error() returns a thread-local message for the last failing call on this thread.
 * Scratch: there is no caller workspace.  Whatever a call needs beyond its arguments (a packed kernel image when the
 * caller did not supply one, dz = dy * act'(y), a channel-padded copy of x for in_q % 4 != 0, host-path staging) is
 * stream-ordered memory from the device's default pool (cudaMallocAsync / cudaFreeAsync on `stream`): no implicit
 * synchronisation, safe under CUDA-graph capture, re-entrant across threads, streams and devices (all library state
 * is per device; the device is the one current when the call is made).
 * Tensors are fp32.  Channel axes are component-BLOCKED: [r(0:C) | i(C:2C) | j(2C:3C) | k(3C:4C)]
 * (reference: complexnn/conv.py:294-307, complexnn/dense.py:131-134, complexnn/utils.py:17-79).
 */
#ifndef QNN_B200_H
#define QNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QNN_ABI_VERSION 2
#if defined(__GNUC__)
#define QNN_API __attribute__((visibility("default")))
#else
#define QNN_API
#endif

typedef enum qnn_status {
    QNN_OK = 0,
    QNN_E_INVALID = -1,     /* bad argument (maps to ValueError in the Python mirror)            */
    QNN_E_UNSUPPORTED = -2, /* shape / option outside what the selected algorithm implements    */
    QNN_E_CUDA = -3,        /* a CUDA runtime / driver call failed                              */
    /* -4 was QNN_E_WORKSPACE in ABI 1 (never returned: the library allocates stream-ordered scratch itself) */
    QNN_E_COMM = -5,        /* NCCL unavailable or a collective failed                          */
    QNN_E_STATE = -6        /* call sequence error (e.g. all-reduce before qnn_comm_init)       */
} qnn_status;

typedef enum qnn_padding { QNN_PAD_VALID = 0, QNN_PAD_SAME = 1, QNN_PAD_CAUSAL = 2 } qnn_padding;

/* keras.activations handled in the fused epilogue (complexnn/conv.py:342-343, dense.py:161-162). */
typedef enum qnn_activation {
    QNN_ACT_LINEAR = 0,
    QNN_ACT_RELU = 1,
    QNN_ACT_TANH = 2,
    QNN_ACT_SIGMOID = 3,
    QNN_ACT_HARD_SIGMOID = 4,
    QNN_ACT_SOFTPLUS = 5,
    QNN_ACT_SOFTSIGN = 6,
    QNN_ACT_ELU = 7,
    QNN_ACT_SELU = 8,
    QNN_ACT_EXPONENTIAL = 9
} qnn_activation;

/* Arithmetic of the contraction.
 * TF32   (fast): operands rounded to nearest tf32, fp32 accumulate in tensor memory (tcgen05.mma kind::tf32);
 *        max|d| / max|ref| ~ 3e-4, a fraction of a percent of outputs miss allclose(rtol 1e-3, atol 1e-3 rms).
 * 3XTF32 (parity-safe): every operand split into hi = rn_tf32(v) and lo = rn_tf32(v - hi), each block is
 *        x_lo.w_hi + x_hi.w_lo + x_hi.w_hi on the tensor cores (three MMAs) -- fp32-faithful like the reference's
 *        arithmetic (complexnn/conv.py:334, dense.py:149); shapes no tensor-core kernel takes run the FP32 kernel.
 *        Forward, data gradient and kernel gradient all have it.
 * FP32:  CUDA-core FMA. */
typedef enum qnn_math { QNN_MATH_TF32 = 0, QNN_MATH_FP32 = 1, QNN_MATH_3XTF32 = 2 } qnn_math;

/* Kernel selection.  AUTO picks the tensor-core kernel when the shape qualifies, else the general kernel. */
typedef enum qnn_algo { QNN_ALGO_AUTO = 0, QNN_ALGO_GENERAL = 1, QNN_ALGO_TENSOR = 2 } qnn_algo;

/* One quaternion convolution (QuaternionConv.__init__ / build, complexnn/conv.py:93-152, 154-286).
 * x:      channels_last  [batch, in_spatial..., 4*in_q]   or channels_first [batch, 4*in_q, in_spatial...]
 * kernel: [kernel[0..rank), in_q, 4*filters]  (the layout qconv_init really produces, complexnn/init.py:91)
 * y:      channels_last  [batch, out_spatial..., 4*filters] or channels_first [batch, 4*filters, out_spatial...]   */
typedef struct qnn_conv_desc {
    int32_t rank;          /* 1, 2 or 3 */
    int32_t batch;
    int32_t in_spatial[3];
    int32_t in_q;          /* quaternion input channels  = real channels / 4 */
    int32_t filters;       /* quaternion output channels = real channels / 4 */
    int32_t kernel[3];
    int32_t stride[3];
    int32_t dilation[3];
    int32_t padding;        /* qnn_padding; CAUSAL only for rank 1 */
    int32_t channels_first; /* 0 = channels_last */
    int32_t activation;     /* qnn_activation */
    int32_t math;           /* qnn_math */
    int32_t algo;           /* qnn_algo */
} qnn_conv_desc;

QNN_API int qnn_abi_version(void);
QNN_API const char* qnn_last_error(void);
/* 1 when a tensor-core kernel will be used for this descriptor's forward, else 0. */
QNN_API int qnn_conv_uses_tensor_cores(const qnn_conv_desc* d);
/* Which forward kernel the descriptor gets (host-only, needs no GPU): a qnn_kernel value, or a negative qnn_status. */
typedef enum qnn_kernel {
    QNN_KERNEL_GENERAL = 0,  /* CUDA cores, fp32: any rank / stride / dilation / layout                                 */
    QNN_KERNEL_TC_ROWS = 1,  /* tcgen05, channels_last rank 1 / dense, resident sub-filters (qnn_hamilton_tc.cu)        */
    QNN_KERNEL_TC_CF = 2,    /* tcgen05, channels_first rank 1 / 2, streamed sub-filters (qnn_hamilton_tc2d.cu)         */
    QNN_KERNEL_SMALL_K = 3   /* CUDA cores, fp32, warp-shuffle tap reuse: in_q < 4 channels_last rank 1 (qnn_smallk.cu)  */
} qnn_kernel;
QNN_API int qnn_conv_forward_kernel(const qnn_conv_desc* d);
QNN_API int qnn_dense_forward_kernel(int64_t rows, int32_t in_q, int32_t q_units, int32_t activation, int32_t math,
                             int32_t algo);
/* How the persistent tensor-core forward kernel of this descriptor spreads its work over the SMs (host-only, needs no GPU;
 * 148 SMs are assumed when no device is present): out4 = { CTAs, whole rounds of one work item (a tile of 128 output
 * positions x one filter tile) per CTA, items of the last partial round, 1 when that round's items are split along the
 * filters into twice as many half-width items }.  Returns the qnn_kernel (out4 stays 0 for the CUDA-core kernels) or a
 * negative qnn_status.  Diagnostics / tests: every item is computed exactly once whatever the split. */
QNN_API int qnn_conv_work_split(const qnn_conv_desc* d, int32_t* out4);
QNN_API int qnn_dense_uses_tensor_cores(int64_t rows, int32_t in_q, int32_t q_units);

/* Which gradients of this layer qnn_*_backward computes on the tensor cores (1) or on the CUDA-core kernels (0) under
 * the descriptor's math / algo (dense: TF32 + AUTO); pointer alignment aside.  Host-only, needs no GPU. */
QNN_API int qnn_conv_backward_uses_tensor_cores(const qnn_conv_desc* d, int32_t* dx_tc, int32_t* dkernel_tc);
QNN_API int qnn_dense_backward_uses_tensor_cores(int64_t rows, int32_t in_q, int32_t q_units, int32_t* dx_tc,
                                         int32_t* dkernel_tc);

/* replaces conv_utils.conv_output_length at complexnn/conv.py:347-372 */
QNN_API int qnn_conv_out_spatial(const qnn_conv_desc* d, int32_t out_spatial[3]);

/* replaces the whole of QuaternionConv.call, complexnn/conv.py:288-345:
 * slice r,i,j,k -> Hamilton expansion -> K.conv{1,2,3}d -> K.bias_add -> activation, fused, nothing expanded in HBM.
 * bias may be NULL (use_bias=False). */
QNN_API int qnn_conv_forward(const qnn_conv_desc* d, const float* x, const float* kernel, const float* bias, float* y,
                     void* stream);

/* replaces QuaternionDense.call, complexnn/dense.py:126-164 (transposed Hamilton table).
 * x [rows, 4*in_q], kernel [in_q, 4*q_units], bias [4*q_units] or NULL, y [rows, 4*q_units]. */
QNN_API int qnn_dense_forward(int64_t rows, int32_t in_q, int32_t q_units, const float* x, const float* kernel,
                      const float* bias, int32_t activation, int32_t math, int32_t algo, float* y, void* stream);

/* Packed kernel images.  The tensor-core kernels consume the stored kernel as a K-major, tf32-rounded core-matrix image
 * (3XTF32: hi and lo parts).  The plain entry points build it in scratch on every call (one extra ~2 us launch); a
 * caller that keeps weights across calls -- the layer mirror does, keyed by a weight version -- packs once per weight
 * update into its own buffer and passes it to the *_packed entry points (one launch per call).
 * kind FORWARD feeds qnn_*_forward_packed, kind DGRAD (the transposed, tap-flipped kernel of the data gradient, read
 * straight from the stored kernel) feeds qnn_*_backward_packed.  qnn_*_packed_bytes returns 0 when the problem has no
 * packed form (it runs on the general kernel); the image depends on the layer (in_q, filters, kernel, dilation, layout),
 * math and algo -- not on batch size or (as long as the same kernel family takes the problem) on the spatial extent.
 * `packed` must be 16-byte aligned.  `kernel` may be NULL in the *_packed calls when a tensor-core kernel runs. */
typedef enum qnn_pack_kind { QNN_PACK_FORWARD = 0, QNN_PACK_DGRAD = 1 } qnn_pack_kind;
QNN_API size_t qnn_conv_packed_bytes(const qnn_conv_desc* d, int32_t kind);
QNN_API int qnn_conv_pack(const qnn_conv_desc* d, int32_t kind, const float* kernel, void* packed, void* stream);
QNN_API int qnn_conv_forward_packed(const qnn_conv_desc* d, const float* x, const float* kernel, const void* packed,
                            const float* bias, float* y, void* stream);
QNN_API size_t qnn_dense_packed_bytes(int64_t rows, int32_t in_q, int32_t q_units, int32_t math, int32_t algo, int32_t kind);
QNN_API int qnn_dense_pack(int64_t rows, int32_t in_q, int32_t q_units, int32_t math, int32_t algo, int32_t kind,
                   const float* kernel, void* packed, void* stream);
QNN_API int qnn_dense_forward_packed(int64_t rows, int32_t in_q, int32_t q_units, const float* x, const float* kernel,
                             const void* packed, const float* bias, int32_t activation, int32_t math, int32_t algo,
                             float* y, void* stream);

/* Gradients TF autodiff derives from the same graph.  `y` is the forward output (needed for the
 * activation derivative; only LINEAR and RELU are differentiable here).  Any of dx / dkernel / dbias may be NULL to
 * skip it.  dkernel / dbias are OVERWRITTEN (not accumulated) and have the stored-kernel / bias shapes, so they can
 * point into a flat gradient bucket that qnn_allreduce_f32 then reduces.
 * math / algo as in the forward: under TF32 / 3XTF32 + AUTO the data gradient of a stride-1 layer runs on the tensor
 * cores (the forward kernel on dz with the transposed, tap-flipped kernel image -- channels_last rank 1 / dense and
 * channels_first rank 1 / 2 / 3); so does the kernel gradient (rank 1 / 2, both layouts; for a relu layer it forms
 * dz = relu'(y) * dy and the bias gradient in the same kernel); FP32 or GENERAL selects the CUDA-core kernels.  The *_packed variants take the caller's cached DGRAD image (or NULL). */
QNN_API int qnn_conv_backward(const qnn_conv_desc* d, const float* x, const float* kernel, const float* y, const float* dy,
                      float* dx, float* dkernel, float* dbias, void* stream);
QNN_API int qnn_dense_backward(int64_t rows, int32_t in_q, int32_t q_units, const float* x, const float* kernel,
                       const float* y, const float* dy, int32_t activation, int32_t math, int32_t algo, float* dx,
                       float* dkernel, float* dbias, void* stream);
QNN_API int qnn_conv_backward_packed(const qnn_conv_desc* d, const float* x, const float* kernel, const void* packed_dgrad,
                             const float* y, const float* dy, float* dx, float* dkernel, float* dbias, void* stream);
QNN_API int qnn_dense_backward_packed(int64_t rows, int32_t in_q, int32_t q_units, const float* x, const float* kernel,
                              const void* packed_dgrad, const float* y, const float* dy, int32_t activation,
                              int32_t math, int32_t algo, float* dx, float* dkernel, float* dbias, void* stream);

/* Host-buffer convenience (the end-to-end call a non-CUDA caller makes): pageable or pinned HOST pointers in, HOST
 * result out; the library stages through stream-ordered device scratch (H2D, kernel and D2H pipelined in chunks over
 * `stream` and two copy streams of its own) and returns after the result landed.  Kernel and bias stay resident on the
 * device between calls, keyed by (host pointer, size, content hash), together with the packed image. */
QNN_API int qnn_conv_forward_host(const qnn_conv_desc* d, const float* x_host, const float* kernel_host,
                          const float* bias_host, float* y_host, void* stream);
QNN_API int qnn_dense_forward_host(int64_t rows, int32_t in_q, int32_t q_units, const float* x_host, const float* kernel_host,
                           const float* bias_host, int32_t activation, int32_t math, int32_t algo, float* y_host,
                           void* stream);

/* Data-parallel gradient exchange (absent from the reference): one NCCL communicator per process.
 * qnn_comm_unique_id fills 128 bytes on rank 0; every rank passes the same bytes to qnn_comm_init. */
QNN_API int qnn_comm_unique_id(void* out_128_bytes);
QNN_API int qnn_comm_init(int32_t rank, int32_t world_size, const void* unique_id_128_bytes);
QNN_API int qnn_allreduce_f32(float* buf, size_t count, void* stream); /* in-place sum */
QNN_API int qnn_comm_destroy(void);

/* Diagnostics: when `device_buffer` (>= 64 * 8 bytes per CTA, i.e. 148 * 512 bytes) is set, the tensor-core kernel
 * records per-CTA clock64() timestamps of its pipeline events into it (layout: tools/tc_trace.py). NULL disables. */
QNN_API int qnn_debug_trace(void* device_buffer, size_t bytes);

/* Number of GPU kernels this library has launched in this process (all threads) -- for bench.py's gpu_launches. */
QNN_API uint64_t qnn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QNN_B200_H */
