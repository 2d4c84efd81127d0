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

// Optional fused "dropout of the output gradient" of the backward norm kernels (vector path): out = dropout(dx) with the mask
// klab_dropout_apply(seed, seed_ptr, p) would draw for a contiguous [rows, d] tensor.
struct NormDrop {
    void* out;
    long long ld;
    float p;
    unsigned long long seed;
    const unsigned long long* seed_ptr;
};


constexpr int WARPS = 4;
constexpr int VEC = 4;            // elements per lane per step on the vector path (8 B of bf16 / 16 B of fp32)
constexpr int MAXV = 8;           // vector path covers d <= 32 * VEC * MAXV = 1024 with d % 128 == 0

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
        const uint2 q = *reinterpret_cast<const uint2*>(p);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
        uint2 q;
        *reinterpret_cast<__nv_bfloat162*>(&q.x) = __floats2bfloat162_rn(v[0], v[1]);
        *reinterpret_cast<__nv_bfloat162*>(&q.y) = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = q;
    }
};

// ---- vector path: the whole row lives in registers (one HBM read), d % 128 == 0, d <= 1024, 8/16-byte aligned rows ----
template <typename T, bool IS_LN, int NV>
__global__ void __launch_bounds__(WARPS * 32)
norm_fwd_vec_kernel(const T* __restrict__ x, RowMap xm, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const T* __restrict__ residual, RowMap rm, T* __restrict__ y, RowMap ym, float* __restrict__ mean_out,
                    float* __restrict__ rstd_out, long long rows, int d, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* xr = x + xm.off(row);
    float v[NV][4];
    float s = 0.0f, ss = 0.0f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        Vec4<T>::load(xr + (k * 32 + lane) * VEC, v[k]);
#pragma unroll
        for (int t = 0; t < 4; ++t) { s += v[k][t]; ss += v[k][t] * v[k][t]; }
    }
    float mean = 0.0f, var;
    if (IS_LN) {
        mean = warp_sum(s) / d;
        float sv = 0.0f;
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
            for (int t = 0; t < 4; ++t) { const float c = v[k][t] - mean; sv += c * c; }
        var = warp_sum(sv) / d;
    } else {
        var = warp_sum(ss) / d;
    }
    const float rstd = rsqrtf(var + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
    T* yr = y + ym.off(row);
    const T* rr = residual ? residual + rm.off(row) : nullptr;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int c0 = (k * 32 + lane) * VEC;
        const float4 g = *reinterpret_cast<const float4*>(gamma + c0);
        float o[4] = {(v[k][0] - mean) * rstd * g.x, (v[k][1] - mean) * rstd * g.y, (v[k][2] - mean) * rstd * g.z, (v[k][3] - mean) * rstd * g.w};
        if (IS_LN) {
            const float4 b = *reinterpret_cast<const float4*>(beta + c0);
            o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
        }
        if (rr) {
            float r4[4];
            Vec4<T>::load(rr + c0, r4);
#pragma unroll
            for (int t = 0; t < 4; ++t) o[t] += r4[t];
        }
        Vec4<T>::store(yr + c0, o);
    }
}

template <typename T, bool IS_LN, int NV>
__global__ void __launch_bounds__(WARPS * 32)
norm_bwd_vec_kernel(const T* __restrict__ dy, RowMap dym, const T* __restrict__ x, RowMap xm, const float* __restrict__ gamma,
                    const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const T* __restrict__ dres, RowMap drm,
                    T* __restrict__ dx, RowMap dxm, float* __restrict__ part_dgamma, float* __restrict__ part_dbeta,
                    long long rows, int d, NormDrop dr) {
    extern __shared__ float sm[];          // [WARPS][d] dgamma (, [WARPS][d] dbeta)
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // optional second output: dx with the dropout mask of the consumer applied (the layer that receives dx as its output
    // gradient starts by multiplying it with its forward mask; doing that here saves one pass over the tensor per sub-layer)
    DropKey dkey = make_drop_key(0, 0.0f);
    if (dr.out) dkey = make_drop_key(dr.seed + (dr.seed_ptr ? *dr.seed_ptr : 0ull), dr.p);
    float g4[NV][4], ag[NV][4], ab[NV][4];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + (k * 32 + lane) * VEC);
        g4[k][0] = g.x; g4[k][1] = g.y; g4[k][2] = g.z; g4[k][3] = g.w;
#pragma unroll
        for (int t = 0; t < 4; ++t) { ag[k][t] = 0.0f; ab[k][t] = 0.0f; }
    }
    for (long long row = static_cast<long long>(blockIdx.x) * WARPS + warp; row < rows;
         row += static_cast<long long>(gridDim.x) * WARPS) {
        const T* dyr = dy + dym.off(row);
        const T* xr = x + xm.off(row);
        const float mean = IS_LN ? mean_in[row] : 0.0f;
        const float rstd = rstd_in[row];
        float dv[NV][4], xh[NV][4];
        float sg = 0.0f, sgx = 0.0f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c0 = (k * 32 + lane) * VEC;
            Vec4<T>::load(dyr + c0, dv[k]);
            Vec4<T>::load(xr + c0, xh[k]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                xh[k][t] = (xh[k][t] - mean) * rstd;
                const float g = dv[k][t] * g4[k][t];
                sg += g;
                sgx += g * xh[k][t];
                ag[k][t] += dv[k][t] * xh[k][t];
                if (IS_LN) ab[k][t] += dv[k][t];
            }
        }
        sg = IS_LN ? warp_sum(sg) / d : 0.0f;
        sgx = warp_sum(sgx) / d;
        T* dxr = dx + dxm.off(row);
        const T* drr = dres ? dres + drm.off(row) : nullptr;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c0 = (k * 32 + lane) * VEC;
            float o[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) o[t] = rstd * (dv[k][t] * g4[k][t] - sg - xh[k][t] * sgx);
            if (drr) {
                float r4[4];
                Vec4<T>::load(drr + c0, r4);
#pragma unroll
                for (int t = 0; t < 4; ++t) o[t] += r4[t];
            }
            Vec4<T>::store(dxr + c0, o);
            if (dr.out) {
                // same arithmetic as klab_dropout_apply on the STORED dx: round to T first, then mask * 1 / (1 - p)
                float q[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) q[t] = to_f32(from_f32<T>(o[t]));
                const uint64_t e0 = static_cast<uint64_t>(row) * static_cast<uint64_t>(d) + static_cast<uint64_t>(c0);    // even
                const uint32_t h0 = drop_hash_pair(dkey, e0 >> 1), h1 = drop_hash_pair(dkey, (e0 >> 1) + 1);
                q[0] *= (h0 & 0xFFFFu) < dkey.thr16 ? dkey.inv_keep : 0.0f;
                q[1] *= (h0 >> 16) < dkey.thr16 ? dkey.inv_keep : 0.0f;
                q[2] *= (h1 & 0xFFFFu) < dkey.thr16 ? dkey.inv_keep : 0.0f;
                q[3] *= (h1 >> 16) < dkey.thr16 ? dkey.inv_keep : 0.0f;
                Vec4<T>::store(reinterpret_cast<T*>(dr.out) + row * dr.ld + c0, q);
            }
        }
    }
    // combine the 4 warps' register partials through shared memory, one row of partials per CTA
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            sm[warp * d + (k * 32 + lane) * VEC + t] = ag[k][t];
            if (IS_LN) sm[(WARPS + warp) * d + (k * 32 + lane) * VEC + t] = ab[k][t];
        }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        float a = 0.0f, b = 0.0f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            a += sm[w * d + c];
            if (IS_LN) b += sm[(WARPS + w) * d + c];
        }
        part_dgamma[static_cast<long long>(blockIdx.x) * d + c] = a;
        if (IS_LN) part_dbeta[static_cast<long long>(blockIdx.x) * d + c] = b;
    }
}

// ---- generic path (any d, any alignment) ----
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

// out[a][c] (+)= sum_p part[a][p, c] for a in {0 (, 1)}: blockDim = (32 cols, 32 part lanes), grid = (ceil(d/32), arrays).
// Launched ~270 times per training step behind every norm backward / bias gradient, so it is sized for latency: 1024 threads
// keep nparts/32 independent loads in flight per thread (it used to be 8 lanes and ~10 us per launch).
constexpr int FIN_LANES = 32;
__global__ void __launch_bounds__(32 * FIN_LANES)
colsum_finalize_kernel(const float* __restrict__ part0, const float* __restrict__ part1, int nparts, int d, float* __restrict__ out0,
                       float* __restrict__ out1, int accumulate) {
    __shared__ float red[FIN_LANES][33];
    const float* part = blockIdx.y ? part1 : part0;
    float* out = blockIdx.y ? out1 : out0;
    const int c = blockIdx.x * 32 + threadIdx.x;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    if (c < d) {
        int p = threadIdx.y;
        for (; p + 3 * FIN_LANES < nparts; p += 4 * FIN_LANES) {
            s0 += part[static_cast<long long>(p) * d + c];
            s1 += part[static_cast<long long>(p + FIN_LANES) * d + c];
            s2 += part[static_cast<long long>(p + 2 * FIN_LANES) * d + c];
            s3 += part[static_cast<long long>(p + 3 * FIN_LANES) * d + c];
        }
        for (; p < nparts; p += FIN_LANES) s0 += part[static_cast<long long>(p) * d + c];
    }
    red[threadIdx.y][threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (threadIdx.y == 0 && c < d) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < FIN_LANES; ++i) t += red[i][threadIdx.x];
        out[c] = accumulate ? out[c] + t : t;
    }
}

// Column sums of a [rows, d] matrix (bias gradients): partial[gridDim.y, d]
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

// bf16 fast path (d % 8 == 0, 16-byte aligned rows): every thread owns 8 adjacent columns and streams rows with 16-byte loads,
// four rows in flight; blockDim = (32 column vectors, 8 row lanes), grid = (ceil(d/256), row_chunks).  HBM/L2-bound.
__global__ void __launch_bounds__(256)
colsum_partial_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, long long ld, long long rows, int d, float* __restrict__ part) {
    __shared__ float red[8][32][9];
    const int c0 = (blockIdx.x * 32 + threadIdx.x) * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    if (c0 < d) {
        const long long step = static_cast<long long>(gridDim.y) * 8;
        long long r = static_cast<long long>(blockIdx.y) * 8 + threadIdx.y;
        const __nv_bfloat16* px = x + c0;
        auto add = [&](const uint4& q) {
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[2 * j] += __uint_as_float(w[j] << 16);
                acc[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
            }
        };
        for (; r + 3 * step < rows; r += 4 * step) {
            const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(px + r * ld));
            const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(px + (r + step) * ld));
            const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(px + (r + 2 * step) * ld));
            const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(px + (r + 3 * step) * ld));
            add(q0); add(q1); add(q2); add(q3);
        }
        for (; r < rows; r += step) add(__ldg(reinterpret_cast<const uint4*>(px + r * ld)));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
    __syncthreads();
    // 256 threads finish the 32 x 8 columns of this block: thread -> (column vector, element)
    const int tid = threadIdx.y * 32 + threadIdx.x, cv = tid >> 3, j = tid & 7;
    const int c = (blockIdx.x * 32 + cv) * 8 + j;
    if (c < d) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][cv][j];
        part[static_cast<long long>(blockIdx.y) * d + c] = t;
    }
}

inline long long colsum_chunks(long long rows, int d, bool vec) {
    const long long col_blocks = vec ? (d + 255) / 256 : (d + 31) / 32;
    long long chunks = (8ll * sm_count() + col_blocks - 1) / col_blocks;       // ~8 CTAs per SM in total
    const long long by_rows = (rows + 31) / 32;                                 // at least 4 rows per row lane
    if (chunks > by_rows) chunks = by_rows;
    if (chunks > 1024) chunks = 1024;
    if (chunks < 1) chunks = 1;
    return chunks;
}

template <typename T>
bool vec_ok(const void* a, long long lda, const void* b, long long ldb, const void* c, long long ldc, const void* e, long long lde,
            const RowMap& m1, const RowMap& m2, int d) {
    if (d % 128 != 0 || d > 32 * VEC * MAXV) return false;
    const uintptr_t mask = sizeof(T) * VEC - 1;
    auto ok = [&](const void* p, long long ld) { return p == nullptr || ((reinterpret_cast<uintptr_t>(p) & mask) == 0 && ld % VEC == 0); };
    return ok(a, lda) && ok(b, ldb) && ok(c, ldc) && ok(e, lde) && m1.group_stride % VEC == 0 && m2.group_stride % VEC == 0;
}

template <typename T, bool IS_LN>
int norm_fwd_t(cudaStream_t st, const void* x, RowMap xm, const float* gamma, const float* beta, const void* res, RowMap rm,
               void* y, RowMap ym, float* mean, float* rstd, long long rows, int d, float eps) {
    const unsigned grid = static_cast<unsigned>((rows + WARPS - 1) / WARPS);
#define KLAB_NORM_FWD_ARGS reinterpret_cast<const T*>(x), xm, gamma, beta, reinterpret_cast<const T*>(res), rm, reinterpret_cast<T*>(y), ym, mean, rstd, rows, d, eps
    if (vec_ok<T>(x, xm.ld, res, rm.ld, y, ym.ld, nullptr, 0, xm, ym, d)) {
        switch (d / 128) {
            case 1: norm_fwd_vec_kernel<T, IS_LN, 1><<<grid, WARPS * 32, 0, st>>>(KLAB_NORM_FWD_ARGS); break;
            case 2: norm_fwd_vec_kernel<T, IS_LN, 2><<<grid, WARPS * 32, 0, st>>>(KLAB_NORM_FWD_ARGS); break;
            case 3: norm_fwd_vec_kernel<T, IS_LN, 3><<<grid, WARPS * 32, 0, st>>>(KLAB_NORM_FWD_ARGS); break;
            case 4: norm_fwd_vec_kernel<T, IS_LN, 4><<<grid, WARPS * 32, 0, st>>>(KLAB_NORM_FWD_ARGS); break;
            case 6: norm_fwd_vec_kernel<T, IS_LN, 6><<<grid, WARPS * 32, 0, st>>>(KLAB_NORM_FWD_ARGS); break;
            case 8: norm_fwd_vec_kernel<T, IS_LN, 8><<<grid, WARPS * 32, 0, st>>>(KLAB_NORM_FWD_ARGS); break;
            default: norm_fwd_kernel<T, IS_LN><<<grid, WARPS * 32, 0, st>>>(KLAB_NORM_FWD_ARGS); break;
        }
    } else {
        norm_fwd_kernel<T, IS_LN><<<grid, WARPS * 32, 0, st>>>(KLAB_NORM_FWD_ARGS);
    }
#undef KLAB_NORM_FWD_ARGS
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

template <typename T, bool IS_LN>
int norm_bwd_t(cudaStream_t st, const void* dy, RowMap dym, const void* x, RowMap xm, const float* gamma, const float* mean,
               const float* rstd, const void* dres, RowMap drm, void* dx, RowMap dxm, float* dgamma, float* dbeta,
               int accumulate, float* workspace, long long rows, int d, NormDrop dr = NormDrop{nullptr, 0, 0.0f, 0ull, nullptr}) {
    int grid = static_cast<int>((rows + WARPS - 1) / WARPS);
    const int cap = sm_count() * 4;
    if (grid > cap) grid = cap;
    float* part_dg = workspace;
    float* part_db = workspace + static_cast<long long>(grid) * d;
#define KLAB_NORM_BWD_ARGS reinterpret_cast<const T*>(dy), dym, reinterpret_cast<const T*>(x), xm, gamma, mean, rstd, reinterpret_cast<const T*>(dres), drm, reinterpret_cast<T*>(dx), dxm, part_dg, part_db, rows, d
    if (dr.out && (dr.ld % VEC != 0 || (reinterpret_cast<uintptr_t>(dr.out) & (sizeof(T) * VEC - 1)) != 0 || dr.ld != d)) return -1;   // caller falls back
    const bool vec = vec_ok<T>(dy, dym.ld, x, xm.ld, dres, drm.ld, dx, dxm.ld, dym, xm, d) && (d / 128 <= 4 || d / 128 == 6 || d / 128 == 8);
    if (vec) {
        const size_t smem = (IS_LN ? 2 : 1) * WARPS * d * sizeof(float);
        switch (d / 128) {
            case 1: norm_bwd_vec_kernel<T, IS_LN, 1><<<grid, WARPS * 32, smem, st>>>(KLAB_NORM_BWD_ARGS, dr); break;
            case 2: norm_bwd_vec_kernel<T, IS_LN, 2><<<grid, WARPS * 32, smem, st>>>(KLAB_NORM_BWD_ARGS, dr); break;
            case 3: norm_bwd_vec_kernel<T, IS_LN, 3><<<grid, WARPS * 32, smem, st>>>(KLAB_NORM_BWD_ARGS, dr); break;
            case 4: norm_bwd_vec_kernel<T, IS_LN, 4><<<grid, WARPS * 32, smem, st>>>(KLAB_NORM_BWD_ARGS, dr); break;
            case 6: norm_bwd_vec_kernel<T, IS_LN, 6><<<grid, WARPS * 32, smem, st>>>(KLAB_NORM_BWD_ARGS, dr); break;
            default: norm_bwd_vec_kernel<T, IS_LN, 8><<<grid, WARPS * 32, smem, st>>>(KLAB_NORM_BWD_ARGS, dr); break;
        }
    } else {
        const size_t smem = (IS_LN ? 2 : 1) * d * sizeof(float);
        if (dr.out) return -1;                                   // generic path has no fused dropout: caller falls back
        norm_bwd_kernel<T, IS_LN><<<grid, WARPS * 32, smem, st>>>(KLAB_NORM_BWD_ARGS);
    }
#undef KLAB_NORM_BWD_ARGS
    KLAB_LAUNCH_CHECK();
    colsum_finalize_kernel<<<dim3((d + 31) / 32, IS_LN ? 2 : 1), dim3(32, FIN_LANES), 0, st>>>(part_dg, part_db, grid, d, dgamma, dbeta, accumulate);
    KLAB_LAUNCH_CHECK();
    count_launch(2);
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

// klab_rmsnorm_bwd that ALSO writes dx_drop = dropout(dx; p, seed (+ *seed_ptr)) -- the mask klab_dropout_apply draws for a
// contiguous [rows, d] tensor -- in the same pass.  dx_drop has row stride d.
int klab_rmsnorm_bwd_dropout(void* stream, int dtype, long long rows, int d, const void* dy, long long lddy, const void* x, long long ldx,
                             const float* gamma, const float* rstd, const void* dres, long long lddres, void* dx, long long lddx,
                             float* dgamma, int accumulate_dgamma, void* workspace, void* dx_drop, float p, unsigned long long seed,
                             const unsigned long long* seed_ptr) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(rows > 0 && d > 0 && dx_drop && p >= 0.0f && p < 1.0f, "rmsnorm_bwd_dropout: bad arguments rows=%lld d=%d p=%f", rows, d, p);
    const RowMap dym{lddy, 0, 0}, xm{ldx, 0, 0}, drm{lddres, 0, 0}, dxm{lddx, 0, 0};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* ws = static_cast<float*>(workspace);
    const NormDrop dr{dx_drop, d, p, seed, seed_ptr};
    int rc = dtype == KLAB_BF16
                 ? norm_bwd_t<__nv_bfloat16, false>(st, dy, dym, x, xm, gamma, nullptr, rstd, dres, drm, dx, dxm, dgamma, nullptr, accumulate_dgamma, ws, rows, d, dr)
                 : norm_bwd_t<float, false>(st, dy, dym, x, xm, gamma, nullptr, rstd, dres, drm, dx, dxm, dgamma, nullptr, accumulate_dgamma, ws, rows, d, dr);
    if (rc != -1) return rc;
    // shapes the vector kernel does not take: two passes
    rc = klab_rmsnorm_bwd(stream, dtype, rows, d, dy, lddy, x, ldx, gamma, rstd, dres, lddres, dx, lddx, dgamma, accumulate_dgamma, workspace);
    if (rc) return rc;
    KLAB_REQUIRE(lddx == d, "rmsnorm_bwd_dropout: the two-pass fallback needs a contiguous dx (lddx=%lld, d=%d)", lddx, d);
    return klab_dropout_apply(stream, dtype, rows * d, dx, dx_drop, p, seed, seed_ptr);
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
    return 1024ll * d * static_cast<long long>(sizeof(float));          // upper bound of colsum_chunks() partial rows
}

// out[c] (+)= sum_r x[r, c]   (bias gradients of the Swin linears, SURVEY.md K5 backward)
int klab_colsum(void* stream, int dtype, long long rows, int d, const void* x, long long ldx, float* out, int accumulate,
                void* workspace) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(rows > 0 && d > 0, "colsum: empty input rows=%lld d=%d", rows, d);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* part = static_cast<float*>(workspace);
    const bool vec = dtype == KLAB_BF16 && d % 8 == 0 && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    const long long chunks = colsum_chunks(rows, d, vec);
    if (vec) {
        colsum_partial_bf16x8_kernel<<<dim3((d + 255) / 256, static_cast<unsigned>(chunks)), dim3(32, 8), 0, st>>>(
            reinterpret_cast<const __nv_bfloat16*>(x), ldx, rows, d, part);
    } else {
        const dim3 grid((d + 31) / 32, static_cast<unsigned>(chunks)), block(32, 8);
        if (dtype == KLAB_BF16)
            colsum_partial_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, rows, d, part);
        else
            colsum_partial_kernel<float><<<grid, block, 0, st>>>(reinterpret_cast<const float*>(x), ldx, rows, d, part);
    }
    KLAB_LAUNCH_CHECK();
    colsum_finalize_kernel<<<dim3((d + 31) / 32, 1), dim3(32, FIN_LANES), 0, st>>>(part, nullptr, static_cast<int>(chunks), d, out, nullptr, accumulate);
    KLAB_LAUNCH_CHECK();
    count_launch(2);
    return KLAB_OK;
}

}  // extern "C"
