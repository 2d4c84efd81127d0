// Bandwidth-bound helpers around the GEMMs: dtype casts, embedding gather / scatter-add (K11 + T8),
// patch-embedding im2col (K1), patch-merging gather / scatter (K7), cross-entropy over materialised logits
// (K10, generic path) and small vector utilities.
#include "gemm.cuh"

namespace klab {
namespace {

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = from_f32<TO>(to_f32(in[i]));
}

// fp32 -> bf16, 8 elements per thread (weights: read 32 B, write 16 B)
__global__ void cast_f32_bf16_vec_kernel(const float4* __restrict__ in, uint4* __restrict__ out, long long n8) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const float4 a = in[2 * i], b = in[2 * i + 1];
        uint4 o;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
        h[0] = __floats2bfloat162_rn(a.x, a.y);
        h[1] = __floats2bfloat162_rn(a.z, a.w);
        h[2] = __floats2bfloat162_rn(b.x, b.y);
        h[3] = __floats2bfloat162_rn(b.z, b.w);
        out[i] = o;
    }
}

// T5 decoder input embedding with _shift_right fused (HF/models/t5/modeling_t5.py:595-614, :682):
//   token(b,t) = t == 0 ? start_id : (labels[b,t-1] == -100 ? pad_id : labels[b,t-1]);  out[b,t,:] = table[token,:]
// shift = 0 gives a plain embedding lookup of ids (frozen text encoder input, decode steps).
// N2 (/root/reference/train.py:55): the arithmetic of the reference's image processor (transformers ViTImageProcessor:
// rescale by 1/255 then (x - mean[c]) / std[c]) on the device instead of the host.  Same rounding sequence as the numpy code it
// replaces: r = float32(double(x) * rescale); y = (r - mean) / std in fp32 with IEEE division (no FMA contraction).
struct ImgNormArgs {
    double rescale;
    float mean[8], std[8];
    long long hw;
    int C, has_norm;
};

template <typename TIn, int V>
__global__ void __launch_bounds__(256) image_normalize_kernel(const TIn* __restrict__ in, float* __restrict__ out, long long nvec, ImgNormArgs a) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long e0 = i * V;
        const int c = static_cast<int>((e0 / a.hw) % a.C);             // hw % V == 0 on the vector path: one channel per vector
        float x[V];
        if constexpr (V == 4) {
            if constexpr (sizeof(TIn) == 1) {
                const uchar4 q = reinterpret_cast<const uchar4*>(in)[i];
                x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
            } else {
                const float4 q = reinterpret_cast<const float4*>(in)[i];
                x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
            }
        } else {
            x[0] = static_cast<float>(in[i]);
        }
        float y[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const float r = __double2float_rn(static_cast<double>(x[k]) * a.rescale);
            y[k] = a.has_norm ? __fdiv_rn(__fsub_rn(r, a.mean[c]), a.std[c]) : r;
        }
        if constexpr (V == 4) reinterpret_cast<float4*>(out)[i] = make_float4(y[0], y[1], y[2], y[3]);
        else out[i] = y[0];
    }
}

template <typename T>
__global__ void embedding_fwd_kernel(const long long* __restrict__ ids, int B, int L, int shift, int start_id, int pad_id,
                                     const T* __restrict__ table, int d, T* __restrict__ out, long long ldo, long long vocab,
                                     int* __restrict__ err) {
    const int row = blockIdx.x;
    const int b = row / L, t = row % L;
    long long tok;
    if (shift) {
        tok = t == 0 ? start_id : ids[static_cast<long long>(b) * L + t - 1];
        if (tok == -100) tok = pad_id;
    } else {
        tok = ids[static_cast<long long>(b) * L + t];
    }
    if (tok < 0 || tok >= vocab) {
        if (threadIdx.x == 0) atomicExch(err, 1);
        tok = 0;
    }
    const T* src = table + tok * d;
    T* dst = out + static_cast<long long>(row) * ldo;
    for (int c = threadIdx.x; c < d; c += blockDim.x) dst[c] = src[c];
}

// dtable[token(b,t), :] += dout[b,t,:]   (fp32 atomics into the tied embedding gradient, K11)
template <typename T>
__global__ void embedding_bwd_kernel(const long long* __restrict__ ids, int B, int L, int shift, int start_id, int pad_id,
                                     const T* __restrict__ dout, long long ldo, int d, float* __restrict__ dtable, long long vocab) {
    const int row = blockIdx.x;
    const int b = row / L, t = row % L;
    long long tok;
    if (shift) {
        tok = t == 0 ? start_id : ids[static_cast<long long>(b) * L + t - 1];
        if (tok == -100) tok = pad_id;
    } else {
        tok = ids[static_cast<long long>(b) * L + t];
    }
    if (tok < 0 || tok >= vocab) return;
    const T* src = dout + static_cast<long long>(row) * ldo;
    float* dst = dtable + tok * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) atomicAdd(&dst[c], to_f32(src[c]));
}

// Patch embedding im2col (Conv2d k = s = P, HF/models/swinv2/modeling_swinv2.py:329):
//   out[(b*Hp + py)*Wp + px, (c*P + ky)*P + kx] = pix[b, c, py*P + ky, px*P + kx]
template <typename T>
__global__ void patchify_kernel(const float* __restrict__ pix, int B, int C, int H, int W, int P, T* __restrict__ out, long long ldo) {
    const int Hp = H / P, Wp = W / P, Kp = C * P * P;
    const long long total = static_cast<long long>(B) * Hp * Wp * Kp;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int kk = static_cast<int>(idx % Kp);
        const long long row = idx / Kp;
        const int px = static_cast<int>(row % Wp), py = static_cast<int>((row / Wp) % Hp), b = static_cast<int>(row / (static_cast<long long>(Wp) * Hp));
        const int kx = kk % P, ky = (kk / P) % P, c = kk / (P * P);
        out[row * ldo + kk] = from_f32<T>(pix[((static_cast<long long>(b) * C + c) * H + py * P + ky) * W + px * P + kx]);
    }
}

// Patch merging gather (HF/models/swinv2/modeling_swinv2.py:374-383): 2x2 neighbourhood, concat order
// (0,0),(1,0),(0,1),(1,1) -> out[(b, y2, x2), q*C + c] = x[(b, 2*y2 + (q & 1), 2*x2 + (q >> 1)), c].
// `scatter` = 1 runs the inverse permutation (backward).
template <typename T>
__global__ void patch_merge_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int res, int C, int scatter) {
    const int r2 = res / 2;
    const long long total = static_cast<long long>(B) * r2 * r2 * 4 * C;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int c = static_cast<int>(idx % C);
        const int qd = static_cast<int>((idx / C) % 4);
        const long long row = idx / (4ll * C);
        const int x2 = static_cast<int>(row % r2), y2 = static_cast<int>((row / r2) % r2), b = static_cast<int>(row / (static_cast<long long>(r2) * r2));
        const int y = 2 * y2 + (qd & 1), x = 2 * x2 + (qd >> 1);
        const long long src = ((static_cast<long long>(b) * res + y) * res + x) * C + c;
        if (scatter) out[src] = in[idx];
        else out[idx] = in[src];
    }
}

// 16-byte variant (C a multiple of 8 bf16 / 4 fp32 elements, 16-byte aligned tensors): one thread moves one 16-byte vector;
// the scalar kernel above ran at a sixth of HBM speed (2-byte accesses, 64-bit divisions per element).
__global__ void __launch_bounds__(256) patch_merge_vec_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int res, int CV,
                                                              int scatter) {
    const int r2 = res / 2;
    const long long total = static_cast<long long>(B) * r2 * r2 * 4 * CV;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int c = static_cast<int>(idx % CV);
        const long long t = idx / CV;
        const int qd = static_cast<int>(t & 3);
        const long long row = t >> 2;
        const int x2 = static_cast<int>(row % r2);
        const long long t2 = row / r2;
        const int y2 = static_cast<int>(t2 % r2), b = static_cast<int>(t2 / r2);
        const int y = 2 * y2 + (qd & 1), x = 2 * x2 + (qd >> 1);
        const long long src = ((static_cast<long long>(b) * res + y) * res + x) * CV + c;
        if (scatter) out[src] = __ldcs(in + idx);
        else out[idx] = __ldg(in + src);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Cross entropy with ignore_index over materialised logits (generic path of K10;
// HF/models/t5/modeling_t5.py:1114-1117): one CTA per row.
//   fwd : lse[row] = logsumexp(logits[row,:]); row_loss = lse - logits[row,label] (0 if ignored)
//   bwd : logits[row,:] <- (softmax - onehot) * (*gscale) / n_valid   (0 for ignored rows), in place
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) ce_fwd_kernel(const T* __restrict__ logits, long long ld, int V, const long long* __restrict__ labels,
                                                     float* __restrict__ lse, float* __restrict__ row_loss, int* __restrict__ err) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const T* lr = logits + row * ld;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < V; c += blockDim.x) mx = fmaxf(mx, to_f32(lr[c]));
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float s = 0.0f;
    for (int c = threadIdx.x; c < V; c += blockDim.x) s += __expf(to_f32(lr[c]) - mx);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
        const float l = mx + __logf(t);
        lse[row] = l;
        const long long lab = labels[row];
        float loss = 0.0f;
        if (lab != -100) {
            if (lab < 0 || lab >= V) atomicExch(err, 2);
            else loss = l - to_f32(lr[lab]);
        }
        row_loss[row] = loss;
    }
}

// Fused LM-head CE, second half: combine the per-tile online-softmax partials of a row (one warp per row, fixed order):
// lse = m + log(sum_i s_i exp(m_i - m)), row_loss = lse - logit[label] (0 for ignored rows)
__global__ void __launch_bounds__(256) lmhead_ce_finalize_kernel(const float2* __restrict__ partials, int num_parts, const float* __restrict__ label_logit,
                                                                 const long long* __restrict__ labels, long long rows, int V,
                                                                 float* __restrict__ lse, float* __restrict__ row_loss, int* __restrict__ err) {
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float2* p = partials + row * num_parts;
    float m = -INFINITY;
    for (int i = lane; i < num_parts; i += 32) m = fmaxf(m, p[i].x);
    m = warp_max(m);
    float s = 0.0f;
    for (int i = lane; i < num_parts; i += 32) {
        const float2 q = p[i];
        if (q.x > -INFINITY) s += q.y * __expf(q.x - m);
    }
    s = warp_sum(s);
    if (lane == 0) {
        const float l = m + __logf(s);
        lse[row] = l;
        const long long lab = labels[row];
        float loss = 0.0f;
        if (lab != -100) {
            if (lab < 0 || lab >= V) atomicExch(err, 2);
            else loss = l - label_logit[row];
        }
        row_loss[row] = loss;
    }
}

// loss = sum(row_loss) / n_valid ; stats[0] = loss, stats[1] = n_valid   (single CTA, deterministic order)
__global__ void __launch_bounds__(256) ce_reduce_kernel(const float* __restrict__ row_loss, const long long* __restrict__ labels,
                                                        long long rows, float* __restrict__ stats) {
    __shared__ float rs[256];
    __shared__ float rc[256];
    float s = 0.0f, c = 0.0f;
    for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
        s += row_loss[r];
        c += labels[r] != -100 ? 1.0f : 0.0f;
    }
    rs[threadIdx.x] = s;
    rc[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            rs[threadIdx.x] += rs[threadIdx.x + o];
            rc[threadIdx.x] += rc[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats[0] = rs[0] / rc[0];          // all-ignored batch -> nan, as torch
        stats[1] = rc[0];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) ce_bwd_kernel(T* __restrict__ logits, long long ld, int V, int ld_pad, const long long* __restrict__ labels,
                                                     const float* __restrict__ lse, const float* __restrict__ stats,
                                                     const float* __restrict__ gscale) {
    const long long row = blockIdx.x;
    T* lr = logits + row * ld;
    const long long lab = labels[row];
    const float g = lab == -100 ? 0.0f : (gscale ? *gscale : 1.0f) / stats[1];
    const float l = lse[row];
    for (int c = threadIdx.x; c < ld_pad; c += blockDim.x) {
        float v = 0.0f;
        if (c < V && lab != -100) v = (__expf(to_f32(lr[c]) - l) - (c == lab ? 1.0f : 0.0f)) * g;
        lr[c] = from_f32<T>(v);
    }
}

// y[i] = x[i] * keep(seed, i) / (1 - p): regenerates the mask a GEMM epilogue applied to element (row*N + col)
template <typename T>
__global__ void dropout_apply_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, unsigned long long seed,
                                     const unsigned long long* __restrict__ seed_ptr, float p) {
    if (seed_ptr) seed += *seed_ptr;
    const DropKey key = make_drop_key(seed, p);
    const long long npairs = (n + 1) / 2;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long pr = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pr < npairs; pr += stride) {
        const uint32_t h = drop_hash_pair(key, static_cast<uint64_t>(pr));
        const long long i = 2 * pr;
        const float m0 = (h & 0xFFFFu) < key.thr16 ? key.inv_keep : 0.0f;
        const float m1 = (h >> 16) < key.thr16 ? key.inv_keep : 0.0f;
        if (i + 1 < n && sizeof(T) == 2) {
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(x + i);
            const float2 f = __bfloat1622float2(v);
            *reinterpret_cast<__nv_bfloat162*>(y + i) = __floats2bfloat162_rn(f.x * m0, f.y * m1);
        } else {
            y[i] = from_f32<T>(to_f32(x[i]) * m0);
            if (i + 1 < n) y[i + 1] = from_f32<T>(to_f32(x[i + 1]) * m1);
        }
    }
}

// Greedy decode step (HF/generation/utils.py:2762-2800): next = argmax(fp32 logits row) (first index on ties);
// finished rows emit pad; ids[b, t] = next; unfinished[b] &= next != eos.
__global__ void __launch_bounds__(256) greedy_step_kernel(const float* __restrict__ logits, long long ld, int V, long long* __restrict__ ids,
                                                          long long ld_ids, int t, int* __restrict__ unfinished, int pad_id, int eos_id) {
    __shared__ float bv[256];
    __shared__ int bi[256];
    const int b = blockIdx.x;
    const float* lr = logits + static_cast<long long>(b) * ld;
    float best = -INFINITY;
    int besti = V;
    if ((V & 3) == 0 && (reinterpret_cast<uintptr_t>(lr) & 15) == 0) {     // 16-byte loads, four in flight per thread
        const float4* lr4 = reinterpret_cast<const float4*>(lr);
        const int n4 = V >> 2;
#pragma unroll 4
        for (int c4 = threadIdx.x; c4 < n4; c4 += blockDim.x) {
            const float4 q = __ldg(lr4 + c4);
            const float vv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)                                  // indices rise within a thread: strict > keeps the first maximum
                if (vv[i] > best) { best = vv[i]; besti = 4 * c4 + i; }
        }
    } else {
        for (int c = threadIdx.x; c < V; c += blockDim.x) {
            const float v = lr[c];
            if (v > best || (v == best && c < besti)) { best = v; besti = c; }
        }
    }
    bv[threadIdx.x] = best;
    bi[threadIdx.x] = besti;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const float v = bv[threadIdx.x + o];
            const int i = bi[threadIdx.x + o];
            if (v > bv[threadIdx.x] || (v == bv[threadIdx.x] && i < bi[threadIdx.x])) { bv[threadIdx.x] = v; bi[threadIdx.x] = i; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int alive = unfinished[b];
        const int nxt = alive ? bi[0] : pad_id;
        ids[static_cast<long long>(b) * ld_ids + t] = nxt;
        unfinished[b] = alive && nxt != eos_id;
    }
}

__global__ void seed_advance_kernel(unsigned long long* c) { *c = (*c) * 6364136223846793005ull + 1442695040888963407ull; }

__global__ void scale_kernel(float* __restrict__ x, long long n, const float* __restrict__ s) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) x[i] *= *s;
}

inline unsigned grid_for(long long n, int block, int per_thread = 1) {
    long long g = (n + static_cast<long long>(block) * per_thread - 1) / (static_cast<long long>(block) * per_thread);
    const long long cap = 32ll * sm_count();
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<unsigned>(g);
}

}  // namespace
}  // namespace klab

using namespace klab;

extern "C" {

int klab_cast(void* stream, int src_dtype, int dst_dtype, long long n, const void* src, void* dst) {
    if (int rc = klab_check_device()) return rc;
    if (n <= 0) return KLAB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (src_dtype == KLAB_F32 && dst_dtype == KLAB_BF16) {
        if (n % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0)
            cast_f32_bf16_vec_kernel<<<grid_for(n / 8, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint4*>(dst), n / 8);
        else
            cast_kernel<float, __nv_bfloat16><<<grid_for(n, 256), 256, 0, st>>>(reinterpret_cast<const float*>(src), reinterpret_cast<__nv_bfloat16*>(dst), n);
    } else if (src_dtype == KLAB_BF16 && dst_dtype == KLAB_F32) {
        cast_kernel<__nv_bfloat16, float><<<grid_for(n, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), reinterpret_cast<float*>(dst), n);
    } else if (src_dtype == KLAB_F32 && dst_dtype == KLAB_F32) {
        cast_kernel<float, float><<<grid_for(n, 256), 256, 0, st>>>(reinterpret_cast<const float*>(src), reinterpret_cast<float*>(dst), n);
    } else {
        cast_kernel<__nv_bfloat16, __nv_bfloat16><<<grid_for(n, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), reinterpret_cast<__nv_bfloat16*>(dst), n);
    }
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

// err_flag: device int, set non-zero if a token id is out of range (checked lazily by the host)
int klab_embedding_fwd(void* stream, int dtype, int B, int L, const long long* ids, int shift_right, int start_id, int pad_id,
                       const void* table, long long vocab, int d, void* out, long long ldo, int* err_flag) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(B > 0 && L > 0, "embedding_fwd: empty input");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == KLAB_BF16)
        embedding_fwd_kernel<__nv_bfloat16><<<B * L, 128, 0, st>>>(ids, B, L, shift_right, start_id, pad_id, reinterpret_cast<const __nv_bfloat16*>(table), d, reinterpret_cast<__nv_bfloat16*>(out), ldo, vocab, err_flag);
    else
        embedding_fwd_kernel<float><<<B * L, 128, 0, st>>>(ids, B, L, shift_right, start_id, pad_id, reinterpret_cast<const float*>(table), d, reinterpret_cast<float*>(out), ldo, vocab, err_flag);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

int klab_embedding_bwd(void* stream, int dtype, int B, int L, const long long* ids, int shift_right, int start_id, int pad_id,
                       const void* dout, long long ldo, int d, float* dtable, long long vocab) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(B > 0 && L > 0, "embedding_bwd: empty input");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == KLAB_BF16)
        embedding_bwd_kernel<__nv_bfloat16><<<B * L, 128, 0, st>>>(ids, B, L, shift_right, start_id, pad_id, reinterpret_cast<const __nv_bfloat16*>(dout), ldo, d, dtable, vocab);
    else
        embedding_bwd_kernel<float><<<B * L, 128, 0, st>>>(ids, B, L, shift_right, start_id, pad_id, reinterpret_cast<const float*>(dout), ldo, d, dtable, vocab);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

int klab_image_normalize(void* stream, int in_is_u8, int B, int C, long long hw, const void* in, double rescale, const float* mean,
                         const float* std, int has_norm, float* out) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(B > 0 && C > 0 && C <= 8 && hw > 0, "image_normalize: bad shape B=%d C=%d hw=%lld", B, C, hw);
    ImgNormArgs a{};
    a.rescale = rescale;
    a.C = C;
    a.hw = hw;
    a.has_norm = has_norm;
    for (int c = 0; c < C; ++c) {
        a.mean[c] = has_norm ? mean[c] : 0.0f;
        a.std[c] = has_norm ? std[c] : 1.0f;
    }
    const long long total = static_cast<long long>(B) * C * hw;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & (in_is_u8 ? 3 : 15)) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec) {
        if (in_is_u8) image_normalize_kernel<uint8_t, 4><<<grid_for(total / 4, 256), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(in), out, total / 4, a);
        else image_normalize_kernel<float, 4><<<grid_for(total / 4, 256), 256, 0, st>>>(reinterpret_cast<const float*>(in), out, total / 4, a);
    } else {
        if (in_is_u8) image_normalize_kernel<uint8_t, 1><<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(in), out, total, a);
        else image_normalize_kernel<float, 1><<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const float*>(in), out, total, a);
    }
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

int klab_patchify(void* stream, int out_dtype, int B, int C, int H, int W, int P, const float* pixels, void* out, long long ldo) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(B > 0 && H % P == 0 && W % P == 0, "patchify: image %dx%d is not a multiple of patch %d", H, W, P);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = static_cast<long long>(B) * C * H * W;
    if (out_dtype == KLAB_BF16)
        patchify_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(pixels, B, C, H, W, P, reinterpret_cast<__nv_bfloat16*>(out), ldo);
    else
        patchify_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(pixels, B, C, H, W, P, reinterpret_cast<float*>(out), ldo);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

int klab_patch_merge(void* stream, int dtype, int B, int res, int C, const void* in, void* out, int scatter) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(B > 0 && res % 2 == 0, "patch_merge: odd grid %d", res);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = static_cast<long long>(B) * res * res * C;
    const int per_vec = dtype == KLAB_BF16 ? 8 : 4;
    if (C % per_vec == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        patch_merge_vec_kernel<<<grid_for(total / per_vec, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), B, res,
                                                                             C / per_vec, scatter);
        KLAB_LAUNCH_CHECK();
        count_launch();
        return KLAB_OK;
    }
    if (dtype == KLAB_BF16)
        patch_merge_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), B, res, C, scatter);
    else
        patch_merge_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out), B, res, C, scatter);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

int klab_greedy_step(void* stream, int B, int V, const float* logits, long long ld, long long* ids, long long ld_ids, int t,
                     int* unfinished, int pad_id, int eos_id) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(B > 0 && V > 0, "greedy_step: empty input");
    greedy_step_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, ld, V, ids, ld_ids, t, unfinished, pad_id, eos_id);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

int klab_seed_advance(void* stream, unsigned long long* seed_counter) {
    if (int rc = klab_check_device()) return rc;
    seed_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(seed_counter);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

int klab_dropout_apply(void* stream, int dtype, long long n, const void* x, void* y, float p, unsigned long long seed,
                       const unsigned long long* seed_ptr) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(n > 0 && p >= 0.0f && p < 1.0f, "dropout_apply: bad arguments n=%lld p=%f", n, p);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    KLAB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 3) == 0, "dropout_apply: pointers must be 4-byte aligned");
    if (dtype == KLAB_BF16)
        dropout_apply_kernel<__nv_bfloat16><<<grid_for(n, 256, 2), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), n, seed, seed_ptr, p);
    else
        dropout_apply_kernel<float><<<grid_for(n, 256, 2), 256, 0, st>>>(reinterpret_cast<const float*>(x), reinterpret_cast<float*>(y), n, seed, seed_ptr, p);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

// stats: device float[2] = {mean loss over non-ignored rows, number of non-ignored rows}
int klab_ce_fwd(void* stream, int dtype, long long rows, int V, const void* logits, long long ld, const long long* labels,
                float* lse, float* row_loss, float* stats, int* err_flag) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(rows > 0 && V > 0, "ce_fwd: empty input");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == KLAB_BF16)
        ce_fwd_kernel<__nv_bfloat16><<<static_cast<unsigned>(rows), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(logits), ld, V, labels, lse, row_loss, err_flag);
    else
        ce_fwd_kernel<float><<<static_cast<unsigned>(rows), 256, 0, st>>>(reinterpret_cast<const float*>(logits), ld, V, labels, lse, row_loss, err_flag);
    KLAB_LAUNCH_CHECK();
    ce_reduce_kernel<<<1, 256, 0, st>>>(row_loss, labels, rows, stats);
    KLAB_LAUNCH_CHECK();
    count_launch(2);
    return KLAB_OK;
}

long long klab_lmhead_ce_workspace_bytes(long long rows, int V) {
    return static_cast<long long>(sizeof(float2)) * rows * klab::lmhead_ce_num_parts(V) + static_cast<long long>(sizeof(float)) * 2 * rows;
}

int klab_lmhead_ce_fwd(void* stream, long long rows, int V, int d, const void* h, long long ldh, const void* E, long long lde, float alpha,
                       const long long* labels, float* lse, float* stats, void* workspace, int* err_flag) {
    using namespace klab;
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(rows > 0 && rows < (1ll << 31) && V > 0 && d > 0 && workspace, "lmhead_ce_fwd: bad arguments rows=%lld V=%d d=%d", rows, V, d);
    KLAB_REQUIRE(((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(E)) & 15) == 0, "lmhead_ce_fwd: operands must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int parts = lmhead_ce_num_parts(V);
    float2* partials = static_cast<float2*>(workspace);
    float* label_logit = reinterpret_cast<float*>(partials + rows * parts);
    float* row_loss = label_logit + rows;
    if (int rc = lmhead_ce_fwd_launch(st, static_cast<int>(rows), V, d, h, ldh, E, lde, alpha, labels, partials, label_logit)) return rc;
    lmhead_ce_finalize_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(partials, parts, label_logit, labels, rows, V, lse, row_loss, err_flag);
    KLAB_LAUNCH_CHECK();
    ce_reduce_kernel<<<1, 256, 0, st>>>(row_loss, labels, rows, stats);
    KLAB_LAUNCH_CHECK();
    count_launch(2);
    return KLAB_OK;
}

int klab_lmhead_ce_bwd_chunk(void* stream, long long rows, int d, const void* h, long long ldh, const void* E_chunk, long long lde, float alpha,
                             const long long* labels, const float* lse, const float* stats, const float* gscale, int v0, int vc,
                             void* dlogits, long long ldd) {
    using namespace klab;
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(rows > 0 && rows < (1ll << 31) && vc > 0 && v0 >= 0 && ldd >= vc, "lmhead_ce_bwd_chunk: bad arguments rows=%lld v0=%d vc=%d ldd=%lld", rows, v0, vc, ldd);
    KLAB_REQUIRE(((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(E_chunk)) & 15) == 0, "lmhead_ce_bwd_chunk: operands must be 16-byte aligned");
    return lmhead_ce_bwd_launch(static_cast<cudaStream_t>(stream), static_cast<int>(rows), vc, d, h, ldh, E_chunk, lde, alpha, labels, lse,
                                stats, gscale, v0, dlogits, ldd);
}

// Overwrites logits[:, 0:ld_pad) with d(loss)/d(logits) * (*gscale); columns [V, ld_pad) are zeroed.
int klab_ce_bwd(void* stream, int dtype, long long rows, int V, void* logits, long long ld, int ld_pad, const long long* labels,
                const float* lse, const float* stats, const float* gscale) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(rows > 0 && V > 0 && ld_pad >= V && ld_pad <= ld, "ce_bwd: bad shape rows=%lld V=%d ld=%lld ld_pad=%d", rows, V, ld, ld_pad);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == KLAB_BF16)
        ce_bwd_kernel<__nv_bfloat16><<<static_cast<unsigned>(rows), 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(logits), ld, V, ld_pad, labels, lse, stats, gscale);
    else
        ce_bwd_kernel<float><<<static_cast<unsigned>(rows), 256, 0, st>>>(reinterpret_cast<float*>(logits), ld, V, ld_pad, labels, lse, stats, gscale);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

}  // extern "C"
