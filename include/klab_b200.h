/*
 * klab_b200.h -- C ABI of libklab_b200.so: hand-written sm_100a kernels for the image-caption
 * training step of Da-Tsuchi/KLab_MultiModalModel (Swin-V2 encoder + T5 encoder/decoder + LM-head CE).
 *
 * The reference has no FFI of its own: its seam is the Python class `MyModel`
 * (/root/reference/models/model.py:8-42) and all arithmetic is delegated to `transformers`
 * (HF/ below = site-packages/transformers 5.5.0).  Each entry point therefore cites the HF function
 * whose eager op chain it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero otherwise (never throws); klab_last_error() gives
 *     the message of the calling thread's last failure;
 *   - all pointers are DEVICE pointers unless stated; tensors are row-major; `ld*` are row strides in
 *     ELEMENTS; `stream` is a cudaStream_t passed as void*;
 *   - dtypes: KLAB_F32 (strict path, fp32 SIMT arithmetic) or KLAB_BF16 (bf16 storage, tcgen05 tensor
 *     cores, fp32 accumulation and statistics);
 *   - the library borrows pointers for the duration of a call and keeps none;
 *   - there is no CPU fallback: on a device that is not sm_100 every compute entry fails with
 *     KLAB_ERR_UNSUPPORTED.
 */
#ifndef KLAB_B200_H
#define KLAB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KLAB_ABI_VERSION 1

enum { KLAB_F32 = 0, KLAB_BF16 = 1 };
enum { KLAB_OK = 0, KLAB_ERR_INVALID = 1, KLAB_ERR_CUDA = 2, KLAB_ERR_UNSUPPORTED = 3 };

/* ---- library / device probes -------------------------------------------------------------- */
int klab_abi_version(void);
const char* klab_last_error(void);
/* 0 iff the current CUDA device is compute capability 10.x (sm_100a code is present in the library). */
int klab_check_device(void);
/* number of kernels this library has launched since load (bench.py's `gpu_launches`). */
long long klab_launch_count(void);

/* ---- K5: GEMM with fused epilogue -----------------------------------------------------------
 * Replaces every nn.Linear on the path: HF/models/swinv2/modeling_swinv2.py:535,578-579,591,385 and
 * HF/models/t5/modeling_t5.py:93-102,178-181,277,298-299,338,1110 (forward, dgrad and wgrad).
 *
 *   D[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
 *
 * A(m,k) is read from A[m*lda + k]  (a_mn_major = 0)  or  A[k*lda + m]  (a_mn_major = 1);
 * B(n,k) is read from B[n*ldb + k]  (b_mn_major = 0)  or  B[k*ldb + n]  (b_mn_major = 1).
 *   forward  y = x W^T      : A = x  [M,K] (0), B = W [N,K] (0)
 *   dgrad    dx = dy W      : A = dy [M,N'](0), B = W [N',K'] read as B(n=k', k=n') (1)
 *   wgrad    dW = dy^T x    : A = dy read as A(m=n', k=row) (1), B = x read as B(n=k', k=row) (1)
 * epilogue, in order:  v += bias[n];  aux_out[m,n] = v;  v = act(v) or v *= act'(aux_in[m,n]);
 *                      v *= dropout_keep(seed, m*N+n)/(1-p);  v += residual[m,n];  v += D_old (accumulate);
 * in_dtype KLAB_BF16 runs the tcgen05/TMEM/TMA kernel (requires lda, ldb multiples of 8 and 16-byte
 * aligned bases); in_dtype KLAB_F32 runs the fp32 SIMT kernel (strict parity path). */
enum { KLAB_ACT_NONE = 0, KLAB_ACT_RELU = 1, KLAB_ACT_GELU = 2, KLAB_ACT_RELU_BWD = 3, KLAB_ACT_GELU_BWD = 4 };

typedef struct klab_gemm_epilogue {
    const float* bias;      /* [N] fp32, or NULL */
    const void* residual;   /* [M,N] ld = ldr, dtype res_dtype, or NULL */
    const void* aux_in;     /* RELU_BWD: saved post-activation; GELU_BWD: saved pre-activation */
    void* aux_out;          /* optional copy of the pre-activation value, dtype out_dtype */
    long long ldr, ld_aux_in, ld_aux_out;
    float alpha;
    int act;
    int accumulate;         /* D += result (D must be initialised) */
    int out_dtype, res_dtype, aux_in_dtype;
    float dropout_p;        /* 0 = off */
    unsigned long long dropout_seed;
} klab_gemm_epilogue;

int klab_gemm(void* stream, int in_dtype, int M, int N, int K,
              const void* A, long long lda, int a_mn_major,
              const void* B, long long ldb, int b_mn_major,
              void* D, long long ldd, const klab_gemm_epilogue* epi);
/* Same contract, forced onto the fp32-accumulating SIMT kernel (on-device cross-check of the tensor-core path). */
int klab_gemm_simt(void* stream, int in_dtype, int M, int N, int K,
                   const void* A, long long lda, int a_mn_major,
                   const void* B, long long ldb, int b_mn_major,
                   void* D, long long ldd, const klab_gemm_epilogue* epi);

#ifdef __cplusplus
}
#endif
#endif /* KLAB_B200_H */
