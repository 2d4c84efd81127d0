// fp32-accumulating SIMT GEMM with the same contract and epilogue as the tcgen05 kernel.
// Two jobs: (1) the strict fp32 parity path (in_dtype KLAB_F32: true fp32 FMAs, so loss and gradients
// can be compared with the fp32 reference at 1e-4); (2) an on-device cross-check of gemm_tc.cu.
// Not a performance path.
#include "gemm.cuh"

namespace klab {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, long long sa_m, long long sa_k, const T* __restrict__ B, long long sb_n,
                 long long sb_k, void* __restrict__ D, long long ldd, int M, int N, int K, klab_gemm_epilogue epi) {
    __shared__ float As[TK][TM + 1];
    __shared__ float Bs[TK][TN + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < K; k0 += TK) {
        for (int idx = threadIdx.x; idx < TM * TK; idx += 256) {
            int mm, kk;
            if (sa_k == 1) { kk = idx % TK; mm = idx / TK; } else { mm = idx % TM; kk = idx / TM; }
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < K) ? to_f32(A[m * sa_m + k * sa_k]) : 0.0f;
        }
        for (int idx = threadIdx.x; idx < TN * TK; idx += 256) {
            int nn, kk;
            if (sb_k == 1) { kk = idx % TK; nn = idx / TK; } else { nn = idx % TN; kk = idx / TN; }
            const int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < N && k < K) ? to_f32(B[n * sb_n + k * sb_k]) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    if (epi.dropout_p > 0.0f && epi.dropout_seed_ptr) epi.dropout_seed += *epi.dropout_seed_ptr;
    const DropKey dr = make_drop_key(epi.dropout_seed, epi.dropout_p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long row = m0 + ty * 4 + i;
        const int col0 = n0 + tx * 4;
        const int nvalid = min(4, N - col0);
        if (row < M && nvalid > 0) epilogue_apply_store<4>(epi, dr, acc[i], row, col0, nvalid, N, D, ldd);
    }
}

}  // namespace

int gemm_simt_launch(cudaStream_t stream, int in_dtype, int M, int N, int K, const void* A, long long lda, int a_mn,
                     const void* B, long long ldb, int b_mn, void* D, long long ldd, const klab_gemm_epilogue& epi) {
    KLAB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    const dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM);
    const long long sa_m = a_mn ? 1 : lda, sa_k = a_mn ? lda : 1;
    const long long sb_n = b_mn ? 1 : ldb, sb_k = b_mn ? ldb : 1;
    if (in_dtype == KLAB_BF16)
        gemm_simt_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(A), sa_m, sa_k,
                                                                  reinterpret_cast<const __nv_bfloat16*>(B), sb_n, sb_k,
                                                                  D, ldd, M, N, K, epi);
    else
        gemm_simt_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(A), sa_m, sa_k,
                                                          reinterpret_cast<const float*>(B), sb_n, sb_k, D, ldd, M, N, K, epi);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

}  // namespace klab
