// GEMM epilogue shared by the tcgen05 kernel (gemm_tc.cu) and the fp32 SIMT kernel (gemm_simt.cu).
#pragma once

#include "../../include/klab_b200.h"
#include "common.cuh"

namespace klab {

// Load `CH` consecutive elements of a row into fp32 registers (vector path when aligned and full).
template <int CH>
__device__ __forceinline__ void load_chunk(const void* base, int dt, long long idx0, int nvalid, float (&out)[CH]) {
    if (dt == KLAB_BF16) {
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + idx0;
        if (CH % 16 == 0 && nvalid == CH && (reinterpret_cast<uintptr_t>(p) & 31) == 0) {
#pragma unroll
            for (int i = 0; i < CH / 16; ++i) {
                uint32_t q[8];
                ld_global_v8(p + i * 16, q);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q[j]));
                    out[i * 16 + 2 * j] = f.x;
                    out[i * 16 + 2 * j + 1] = f.y;
                }
            }
        } else if (CH % 8 == 0 && nvalid == CH && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < CH / 8; ++i) {
                const uint4 q = reinterpret_cast<const uint4*>(p)[i];
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(h[j]);
                    out[i * 8 + 2 * j] = f.x;
                    out[i * 8 + 2 * j + 1] = f.y;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < CH; ++i) out[i] = i < nvalid ? __bfloat162float(p[i]) : 0.0f;
        }
    } else {
        const float* p = reinterpret_cast<const float*>(base) + idx0;
        if (CH % 8 == 0 && nvalid == CH && (reinterpret_cast<uintptr_t>(p) & 31) == 0) {
#pragma unroll
            for (int i = 0; i < CH / 8; ++i) {
                uint32_t q[8];
                ld_global_v8(p + i * 8, q);
#pragma unroll
                for (int j = 0; j < 8; ++j) out[i * 8 + j] = __uint_as_float(q[j]);
            }
        } else if (nvalid == CH && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) {
                const float4 q = reinterpret_cast<const float4*>(p)[i];
                out[i * 4] = q.x; out[i * 4 + 1] = q.y; out[i * 4 + 2] = q.z; out[i * 4 + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < CH; ++i) out[i] = i < nvalid ? p[i] : 0.0f;
        }
    }
}

template <int CH>
__device__ __forceinline__ void store_chunk(void* base, int dt, long long idx0, int nvalid, const float (&v)[CH]) {
    if (dt == KLAB_BF16) {
        __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base) + idx0;
        if (CH % 16 == 0 && nvalid == CH && (reinterpret_cast<uintptr_t>(p) & 31) == 0) {
#pragma unroll
            for (int i = 0; i < CH / 16; ++i) {
                uint32_t q[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(v[i * 16 + 2 * j], v[i * 16 + 2 * j + 1]);
                    q[j] = *reinterpret_cast<const uint32_t*>(&h);
                }
                st_global_v8(p + i * 16, q);
            }
        } else if (CH % 8 == 0 && nvalid == CH && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < CH / 8; ++i) {
                uint4 q;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[i * 8 + 2 * j], v[i * 8 + 2 * j + 1]);
                reinterpret_cast<uint4*>(p)[i] = q;
            }
        } else {
#pragma unroll
            for (int i = 0; i < CH; ++i)
                if (i < nvalid) p[i] = __float2bfloat16_rn(v[i]);
        }
    } else {
        float* p = reinterpret_cast<float*>(base) + idx0;
        if (CH % 8 == 0 && nvalid == CH && (reinterpret_cast<uintptr_t>(p) & 31) == 0) {
#pragma unroll
            for (int i = 0; i < CH / 8; ++i) {
                uint32_t q[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) q[j] = __float_as_uint(v[i * 8 + j]);
                st_global_v8(p + i * 8, q);
            }
        } else if (nvalid == CH && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < CH / 4; ++i)
                reinterpret_cast<float4*>(p)[i] = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < CH; ++i)
                if (i < nvalid) p[i] = v[i];
        }
    }
}

// Apply the epilogue to CH consecutive columns [col0, col0+nvalid) of output row `row` and store.
// `aux_raw` (optional): the CH bf16 values of e.aux_in for this chunk, already fetched by the caller (software prefetch: the
// tcgen05 kernel loads the next chunk's aux values while it works on the current one -- the dependent global load used to be
// the largest single stall of the activation-backward epilogues).
template <int CH>
__device__ __forceinline__ void epilogue_apply_store(const klab_gemm_epilogue& e, const DropKey& dr, float (&v)[CH],
                                                     long long row, long long col0, int nvalid, int N,
                                                     void* D, long long ldd, const uint32_t* aux_raw = nullptr) {
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] *= e.alpha;
    if (e.bias) {
        const float* bp = e.bias + col0;
        if (CH % 4 == 0 && nvalid == CH && (reinterpret_cast<uintptr_t>(bp) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + i);
                v[i * 4] += b4.x; v[i * 4 + 1] += b4.y; v[i * 4 + 2] += b4.z; v[i * 4 + 3] += b4.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < CH; ++i)
                if (i < nvalid) v[i] += __ldg(bp + i);
        }
    }
    if (e.act == KLAB_ACT_GELU_SAVE_GRAD) {
        float gp[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) gelu_and_grad(v[i], v[i], gp[i]);
        if (e.aux_out) store_chunk<CH>(e.aux_out, e.out_dtype, row * e.ld_aux_out + col0, nvalid, gp);
    } else if (e.aux_out) {
        store_chunk<CH>(e.aux_out, e.out_dtype, row * e.ld_aux_out + col0, nvalid, v);
    }
    if (e.act == KLAB_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = fmaxf(v[i], 0.0f);
    } else if (e.act == KLAB_ACT_GELU) {
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = gelu_erf(v[i]);
    } else if (e.act == KLAB_ACT_RELU_BWD || e.act == KLAB_ACT_GELU_BWD || e.act == KLAB_ACT_MUL_AUX) {
        float a[CH];
        if (aux_raw) {
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) {
                a[2 * j] = __uint_as_float(aux_raw[j] << 16);
                a[2 * j + 1] = __uint_as_float(aux_raw[j] & 0xffff0000u);
            }
        } else {
            load_chunk<CH>(e.aux_in, e.aux_in_dtype, row * e.ld_aux_in + col0, nvalid, a);
        }
        if (e.act == KLAB_ACT_RELU_BWD) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = a[i] > 0.0f ? v[i] : 0.0f;
        } else if (e.act == KLAB_ACT_MUL_AUX) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] *= a[i];
        } else {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] *= gelu_erf_grad(a[i]);
        }
    }
    if (dr.on) {
        const uint64_t base = static_cast<uint64_t>(row) * static_cast<uint64_t>(N) + static_cast<uint64_t>(col0);
        dropout_apply_run<CH>(dr, base, v);
    }
    if (e.residual) {
        float r[CH];
        load_chunk<CH>(e.residual, e.res_dtype, row * e.ldr + col0, nvalid, r);
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] += r[i];
    }
    if (e.accumulate) {
        float r[CH];
        load_chunk<CH>(D, e.out_dtype, row * ldd + col0, nvalid, r);
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] += r[i];
    }
    store_chunk<CH>(D, e.out_dtype, row * ldd + col0, nvalid, v);
}

// launchers (defined in gemm_tc.cu / gemm_simt.cu)
int gemm_tc_launch(cudaStream_t stream, int M, int N, int K, const void* A, long long lda, int a_mn,
                   const void* B, long long ldb, int b_mn, void* D, long long ldd, const klab_gemm_epilogue& epi);
int gemm_simt_launch(cudaStream_t stream, int in_dtype, int M, int N, int K, const void* A, long long lda, int a_mn,
                     const void* B, long long ldb, int b_mn, void* D, long long ldd, const klab_gemm_epilogue& epi);

void gemm_set_force(int cta2, int bn, int splits);
void gemm_last_config(int* bn, int* splits, int* cta2);

int lmhead_ce_num_parts(int V);
int lmhead_ce_fwd_launch(cudaStream_t stream, int M, int V, int d, const void* h, long long ldh, const void* E, long long lde, float alpha,
                         const long long* labels, float2* partials, float* label_logit);
int lmhead_ce_bwd_launch(cudaStream_t stream, int M, int vc, int d, const void* h, long long ldh, const void* E_chunk, long long lde,
                         float alpha, const long long* labels, const float* lse, const float* stats, const float* gscale, int v0,
                         void* dlogits, long long ldd);

void count_launch(int n = 1);

}  // namespace klab
