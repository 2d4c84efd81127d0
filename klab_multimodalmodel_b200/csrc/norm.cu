// K6 / K8: fused LayerNorm (Swin-V2 res-post-norm form) and T5 RMSNorm, forward and backward.
//
//   RMSNorm  (HF/models/t5/modeling_t5.py:55-68):      y = x * rsqrt(mean(x^2) + eps) * gamma
//   LayerNorm (HF/models/swinv2/modeling_swinv2.py:273,386,707-712,969):
//                                                      y = (x - mean) * rsqrt(var + eps) * gamma + beta  (+ residual)
// Statistics and reductions are fp32 regardless of the storage dtype.  One warp per row; the second pass over
// the row hits L1, so HBM traffic is one read of x (+ residual) and one write of y.
//
// Row addressing supports "grouped" tensors so that the concat of [image tokens; text tokens]
// (/root/reference/models/model.py:23) costs no copy: row r of a tensor lives at
//   base + (r / rows_per_group) * group_stride + (r % rows_per_group) * ld.
#include "common.cuh"

namespace klab {
void count_launch(int n = 1);
namespace {

struct RowMap {
    long long ld;
    long long group_stride;
    int rows_per_group;
    __device__ __forceinline__ long long off(long long r) const {
        return rows_per_group > 0 ? (r / rows_per_group) * group_stride + (r % rows_per_group) * ld : r * ld;
    }
};

constexpr int WARPS = 4;

template <typename T, bool IS_LN>
__global__ void __launch_bounds__(WARPS * 32)
norm_fwd_kernel(const T* __restrict__ x, RowMap xm, const float* __restrict__ gamma, const float* __restrict__ beta,
                const T* __restrict__ residual, RowMap rm, T* __restrict__ y, RowMap ym, float* __restrict__ mean_out,
                float* __restrict__ rstd_out, long long rows, int d, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* xr = x + xm.off(row);
    float s = 0.0f, ss = 0.0f;
    for (int c = lane; c < d; c += 32) {
        const float v = to_f32(xr[c]);
        s += v;
        ss += v * v;
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    float mean = 0.0f, var;
    if (IS_LN) {
        mean = s / d;
        // two-pass variance for accuracy (matches aten native_layer_norm to fp32 rounding)
        float sv = 0.0f;
        for (int c = lane; c < d; c += 32) {
            const float v = to_f32(xr[c]) - mean;
            sv += v * v;
        }
        var = warp_sum(sv) / d;
    } else {
        var = ss / d;
    }
    const float rstd = rsqrtf(var + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
    T* yr = y + ym.off(row);
    const T* rr = residual ? residual + rm.off(row) : nullptr;
    for (int c = lane; c < d; c += 32) {
        float v = (to_f32(xr[c]) - mean) * rstd * gamma[c];
        if (IS_LN) v += beta[c];
        if (rr) v += to_f32(rr[c]);
        yr[c] = from_f32<T>(v);
    }
}

// dx = rstd * (g - mean(g) [LN only] - xhat * mean(g * xhat)) (+ dres),   g = dy * gamma
// partial dgamma / dbeta: per-CTA column sums written to workspace [gridDim.x, d] (reduced by colsum_finalize).
template <typename T, bool IS_LN>
__global__ void __launch_bounds__(WARPS * 32)
norm_bwd_kernel(const T* __restrict__ dy, RowMap dym, const T* __restrict__ x, RowMap xm, const float* __restrict__ gamma,
                const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const T* __restrict__ dres, RowMap drm,
                T* __restrict__ dx, RowMap dxm, float* __restrict__ part_dgamma, float* __restrict__ part_dbeta,
                long long rows, int d) {
    extern __shared__ float sm[];          // [d] dgamma, [d] dbeta
    float* s_dg = sm;
    float* s_db = sm + d;
    for (int c = threadIdx.x; c < (IS_LN ? 2 * d : d); c += blockDim.x) sm[c] = 0.0f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    for (long long row = static_cast<long long>(blockIdx.x) * WARPS + warp; row < rows;
         row += static_cast<long long>(gridDim.x) * WARPS) {
        const T* dyr = dy + dym.off(row);
        const T* xr = x + xm.off(row);
        const float mean = IS_LN ? mean_in[row] : 0.0f;
        const float rstd = rstd_in[row];
        float sg = 0.0f, sgx = 0.0f;
        for (int c = lane; c < d; c += 32) {
            const float g = to_f32(dyr[c]) * gamma[c];
            const float xh = (to_f32(xr[c]) - mean) * rstd;
            sg += g;
            sgx += g * xh;
        }
        sg = IS_LN ? warp_sum(sg) / d : 0.0f;
        sgx = warp_sum(sgx) / d;
        T* dxr = dx + dxm.off(row);
        const T* drr = dres ? dres + drm.off(row) : nullptr;
        for (int c = lane; c < d; c += 32) {
            const float dyv = to_f32(dyr[c]);
            const float xh = (to_f32(xr[c]) - mean) * rstd;
            float v = rstd * (dyv * gamma[c] - sg - xh * sgx);
            if (drr) v += to_f32(drr[c]);
            dxr[c] = from_f32<T>(v);
            atomicAdd(&s_dg[c], dyv * xh);
            if (IS_LN) atomicAdd(&s_db[c], dyv);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        part_dgamma[static_cast<long long>(blockIdx.x) * d + c] = s_dg[c];
        if (IS_LN) part_dbeta[static_cast<long long>(blockIdx.x) * d + c] = s_db[c];
    }
}

// out[c] (+)= sum_p part[p, c]
__global__ void colsum_finalize_kernel(const float* __restrict__ part, int nparts, int d, float* __restrict__ out, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    float s = 0.0f;
    for (int p = 0; p < nparts; ++p) s += part[static_cast<long long>(p) * d + c];
    out[c] = accumulate ? out[c] + s : s;
}

// Column sums of a [rows, d] matrix (bias gradients): partial[gridDim.x, d]
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ x, long long ld, long long rows, int d, float* __restrict__ part) {
    // blockDim = (32 columns, 8 row lanes); grid = (ceil(d/32), row_chunks)
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float s = 0.0f;
    if (c < d)
        for (long long r = static_cast<long long>(blockIdx.y) * 8 + threadIdx.y; r < rows; r += static_cast<long long>(gridDim.y) * 8)
            s += to_f32(x[r * ld + c]);
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < d) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        part[static_cast<long long>(blockIdx.y) * d + c] = t;
    }
}

template <typename T, bool IS_LN>
int norm_fwd_t(cudaStream_t st, const void* x, RowMap xm, const float* gamma, const float* beta, const void* res, RowMap rm,
               void* y, RowMap ym, float* mean, float* rstd, long long rows, int d, float eps) {
    const unsigned grid = static_cast<unsigned>((rows + WARPS - 1) / WARPS);
    norm_fwd_kernel<T, IS_LN><<<grid, WARPS * 32, 0, st>>>(reinterpret_cast<const T*>(x), xm, gamma, beta,
                                                           reinterpret_cast<const T*>(res), rm, reinterpret_cast<T*>(y), ym,
                                                           mean, rstd, rows, d, eps);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

template <typename T, bool IS_LN>
int norm_bwd_t(cudaStream_t st, const void* dy, RowMap dym, const void* x, RowMap xm, const float* gamma, const float* mean,
               const float* rstd, const void* dres, RowMap drm, void* dx, RowMap dxm, float* dgamma, float* dbeta,
               int accumulate, float* workspace, long long rows, int d) {
    int grid = static_cast<int>((rows + WARPS - 1) / WARPS);
    const int cap = sm_count() * 4;
    if (grid > cap) grid = cap;
    float* part_dg = workspace;
    float* part_db = workspace + static_cast<long long>(grid) * d;
    const size_t smem = (IS_LN ? 2 : 1) * d * sizeof(float);
    norm_bwd_kernel<T, IS_LN><<<grid, WARPS * 32, smem, st>>>(
        reinterpret_cast<const T*>(dy), dym, reinterpret_cast<const T*>(x), xm, gamma, mean, rstd,
        reinterpret_cast<const T*>(dres), drm, reinterpret_cast<T*>(dx), dxm, part_dg, part_db, rows, d);
    KLAB_LAUNCH_CHECK();
    colsum_finalize_kernel<<<(d + 127) / 128, 128, 0, st>>>(part_dg, grid, d, dgamma, accumulate);
    KLAB_LAUNCH_CHECK();
    if (IS_LN) {
        colsum_finalize_kernel<<<(d + 127) / 128, 128, 0, st>>>(part_db, grid, d, dbeta, accumulate);
        KLAB_LAUNCH_CHECK();
    }
    count_launch(IS_LN ? 3 : 2);
    return KLAB_OK;
}

}  // namespace
}  // namespace klab

using namespace klab;

extern "C" {

long long klab_norm_bwd_workspace_bytes(long long rows, int d) {
    long long grid = (rows + WARPS - 1) / WARPS;
    const long long cap = 4ll * sm_count();
    if (grid > cap) grid = cap;
    return 2 * grid * d * static_cast<long long>(sizeof(float));
}

int klab_rmsnorm_fwd(void* stream, int dtype, long long rows, int d, const void* x, long long ldx, const float* gamma, float eps,
                     void* y, long long ldy, int y_rows_per_group, long long y_group_stride, float* rstd_out) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(rows > 0 && d > 0, "rmsnorm_fwd: empty input rows=%lld d=%d", rows, d);
    const RowMap xm{ldx, 0, 0}, ym{ldy, y_group_stride, y_rows_per_group}, none{0, 0, 0};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dtype == KLAB_BF16
               ? norm_fwd_t<__nv_bfloat16, false>(st, x, xm, gamma, nullptr, nullptr, none, y, ym, nullptr, rstd_out, rows, d, eps)
               : norm_fwd_t<float, false>(st, x, xm, gamma, nullptr, nullptr, none, y, ym, nullptr, rstd_out, rows, d, eps);
}

int klab_rmsnorm_bwd(void* stream, int dtype, long long rows, int d, const void* dy, long long lddy, const void* x, long long ldx,
                     const float* gamma, const float* rstd, const void* dres, long long lddres, void* dx, long long lddx,
                     float* dgamma, int accumulate_dgamma, void* workspace) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(rows > 0 && d > 0, "rmsnorm_bwd: empty input rows=%lld d=%d", rows, d);
    const RowMap dym{lddy, 0, 0}, xm{ldx, 0, 0}, drm{lddres, 0, 0}, dxm{lddx, 0, 0};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* ws = static_cast<float*>(workspace);
    return dtype == KLAB_BF16
               ? norm_bwd_t<__nv_bfloat16, false>(st, dy, dym, x, xm, gamma, nullptr, rstd, dres, drm, dx, dxm, dgamma, nullptr, accumulate_dgamma, ws, rows, d)
               : norm_bwd_t<float, false>(st, dy, dym, x, xm, gamma, nullptr, rstd, dres, drm, dx, dxm, dgamma, nullptr, accumulate_dgamma, ws, rows, d);
}

int klab_layernorm_fwd(void* stream, int dtype, long long rows, int d, const void* x, long long ldx, const float* gamma,
                       const float* beta, float eps, const void* residual, long long ldres, void* y, long long ldy,
                       int y_rows_per_group, long long y_group_stride, float* mean_out, float* rstd_out) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(rows > 0 && d > 0, "layernorm_fwd: empty input rows=%lld d=%d", rows, d);
    const RowMap xm{ldx, 0, 0}, rm{ldres, 0, 0}, ym{ldy, y_group_stride, y_rows_per_group};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dtype == KLAB_BF16
               ? norm_fwd_t<__nv_bfloat16, true>(st, x, xm, gamma, beta, residual, rm, y, ym, mean_out, rstd_out, rows, d, eps)
               : norm_fwd_t<float, true>(st, x, xm, gamma, beta, residual, rm, y, ym, mean_out, rstd_out, rows, d, eps);
}

int klab_layernorm_bwd(void* stream, int dtype, long long rows, int d, const void* dy, long long lddy, int dy_rows_per_group,
                       long long dy_group_stride, const void* x, long long ldx, const float* gamma, const float* mean,
                       const float* rstd, const void* dres, long long lddres, void* dx, long long lddx, float* dgamma,
                       float* dbeta, int accumulate_dparams, void* workspace) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(rows > 0 && d > 0, "layernorm_bwd: empty input rows=%lld d=%d", rows, d);
    const RowMap dym{lddy, dy_group_stride, dy_rows_per_group}, xm{ldx, 0, 0}, drm{lddres, 0, 0}, dxm{lddx, 0, 0};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* ws = static_cast<float*>(workspace);
    return dtype == KLAB_BF16
               ? norm_bwd_t<__nv_bfloat16, true>(st, dy, dym, x, xm, gamma, mean, rstd, dres, drm, dx, dxm, dgamma, dbeta, accumulate_dparams, ws, rows, d)
               : norm_bwd_t<float, true>(st, dy, dym, x, xm, gamma, mean, rstd, dres, drm, dx, dxm, dgamma, dbeta, accumulate_dparams, ws, rows, d);
}

long long klab_colsum_workspace_bytes(long long rows, int d) {
    long long chunks = (rows + 255) / 256;
    if (chunks > 64) chunks = 64;
    if (chunks < 1) chunks = 1;
    return chunks * d * static_cast<long long>(sizeof(float));
}

// out[c] (+)= sum_r x[r, c]   (bias gradients of the Swin linears, SURVEY.md K5 backward)
int klab_colsum(void* stream, int dtype, long long rows, int d, const void* x, long long ldx, float* out, int accumulate,
                void* workspace) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(rows > 0 && d > 0, "colsum: empty input rows=%lld d=%d", rows, d);
    long long chunks = (rows + 255) / 256;
    if (chunks > 64) chunks = 64;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid((d + 31) / 32, static_cast<unsigned>(chunks)), block(32, 8);
    float* part = static_cast<float*>(workspace);
    if (dtype == KLAB_BF16)
        colsum_partial_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, rows, d, part);
    else
        colsum_partial_kernel<float><<<grid, block, 0, st>>>(reinterpret_cast<const float*>(x), ldx, rows, d, part);
    KLAB_LAUNCH_CHECK();
    colsum_finalize_kernel<<<(d + 127) / 128, 128, 0, st>>>(part, static_cast<int>(chunks), d, out, accumulate);
    KLAB_LAUNCH_CHECK();
    count_launch(2);
    return KLAB_OK;
}

}  // extern "C"
