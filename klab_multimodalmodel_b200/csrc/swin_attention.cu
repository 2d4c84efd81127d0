// K2 + K3 + K4 (generic path): Swin-V2 shifted-window cosine attention, forward and backward, plus the
// continuous-position-bias (CPB) MLP.  Follows HF/models/swinv2/modeling_swinv2.py:
//   * window_partition / cyclic roll / window_reverse (:146-166, :678, :698) are folded into index math:
//     q/k/v/ctx stay in the natural [B*H*W, C] token order that the projection GEMMs use, and the kernel
//     gathers the tokens of window (wy, wx) from ((wy*w + ty + shift) % H, (wx*w + tx + shift) % W);
//   * the shift mask (:627-653, built on the CPU and copied every call in the reference) is computed from the
//     coordinates: region(y) = 0 | 1 | 2 for y < H-w | y < H-shift | else; tokens of different regions get -200
//     (the reference adds the -100 mask twice, :465-468);
//   * cosine attention (:444-449): S = normalize(q) normalize(k)^T * exp(min(logit_scale, ln 100));
//   * bias (:450-460): 16 * sigmoid(MLP(coords)[index]) precomputed once per block per step by cpb kernels.
// CUDA-core fp32 arithmetic, any window size / head dim; the bf16 hot path is swin_attention_tc.cu.
#include <cstdlib>

#include "common.cuh"

namespace klab {
void count_launch(int n = 1);
namespace {

constexpr int NW = 4;
constexpr float LOGIT_MAX = 4.605170185988092f;      // ln(1/0.01)
constexpr float NORM_EPS = 1e-12f;                   // F.normalize eps

struct SwinArgs {
    const void *q, *k, *v, *ctx, *dctx;
    void *out, *dq, *dk, *dv;
    long long ld;              // row stride of q/k/v and dq/dk/dv (elements)
    long long ldc;             // row stride of ctx / dctx
    int B, res, heads, d, w, shift;
    const float* logit_scale;  // [heads]
    const float* bias;         // [heads, N, N]  (16*sigmoid(cpb))
    float* lse;                // [B*nW, heads, N]
    float* dbias;              // [heads, N, N]
    float* dlogit_scale;       // [heads]
    int nchunks;
};

__device__ __forceinline__ int region_of(int y, int res, int w, int shift) { return y < res - w ? 0 : (y < res - shift ? 1 : 2); }

// token index (row in [B*res*res]) and mask region of in-window token n of window `win` of image b
__device__ __forceinline__ void window_token(const SwinArgs& a, int b, int win, int n, int& tok, int& region) {
    const int nwx = a.res / a.w;
    const int wy = win / nwx, wx = win % nwx;
    const int ys = wy * a.w + n / a.w, xs = wx * a.w + n % a.w;        // coordinates in the shifted frame
    int y = ys + a.shift, x = xs + a.shift;
    if (y >= a.res) y -= a.res;
    if (x >= a.res) x -= a.res;
    tok = (b * a.res + y) * a.res + x;
    region = a.shift > 0 ? region_of(ys, a.res, a.w, a.shift) * 3 + region_of(xs, a.res, a.w, a.shift) : 0;
}

template <typename T>
__global__ void __launch_bounds__(NW * 32) swin_attn_fwd_kernel(SwinArgs a) {
    extern __shared__ float sm[];
    const int N = a.w * a.w, d = a.d, dp = d + 1;
    float* qs = sm;
    float* ks = qs + N * dp;
    float* vs = ks + N * dp;
    float* pbuf = vs + N * dp;                        // [NW][N]
    int* tok = reinterpret_cast<int*>(pbuf + NW * N); // [N]
    int* reg = tok + N;                               // [N]
    const int nW = (a.res / a.w) * (a.res / a.w);
    const int bw = blockIdx.x, b = bw / nW, win = bw % nW, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = threadIdx.x; n < N; n += blockDim.x) window_token(a, b, win, n, tok[n], reg[n]);
    __syncthreads();
    // load + L2-normalise q, k rows (one warp per row)
    for (int n = warp; n < N; n += NW) {
        const long long base = static_cast<long long>(tok[n]) * a.ld + h * d;
        float sq = 0.0f, sk = 0.0f;
        for (int c = lane; c < d; c += 32) {
            const float qv = to_f32(reinterpret_cast<const T*>(a.q)[base + c]);
            const float kv = to_f32(reinterpret_cast<const T*>(a.k)[base + c]);
            qs[n * dp + c] = qv;
            ks[n * dp + c] = kv;
            vs[n * dp + c] = to_f32(reinterpret_cast<const T*>(a.v)[base + c]);
            sq += qv * qv;
            sk += kv * kv;
        }
        const float iq = 1.0f / fmaxf(sqrtf(warp_sum(sq)), NORM_EPS);
        const float ik = 1.0f / fmaxf(sqrtf(warp_sum(sk)), NORM_EPS);
        for (int c = lane; c < d; c += 32) {
            qs[n * dp + c] *= iq;
            ks[n * dp + c] *= ik;
        }
    }
    __syncthreads();
    const float scale = __expf(fminf(a.logit_scale[h], LOGIT_MAX));
    const float* bias = a.bias + static_cast<long long>(h) * N * N;
    float* pw = pbuf + warp * N;
    for (int i = warp; i < N; i += NW) {
        float mx = -INFINITY;
        for (int j = lane; j < N; j += 32) {
            float s = 0.0f;
            for (int c = 0; c < d; ++c) s = fmaf(qs[i * dp + c], ks[j * dp + c], s);
            s = s * scale + bias[i * N + j];
            if (reg[i] != reg[j]) s += -200.0f;
            pw[j] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int j = lane; j < N; j += 32) {
            const float e = __expf(pw[j] - mx);
            pw[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        __syncwarp();
        if (lane == 0) a.lse[(static_cast<long long>(bw) * a.heads + h) * N + i] = mx + __logf(sum);
        T* op = reinterpret_cast<T*>(a.out) + static_cast<long long>(tok[i]) * a.ldc + h * d;
        for (int c = lane; c < d; c += 32) {
            float acc = 0.0f;
            for (int j = 0; j < N; ++j) acc = fmaf(pw[j], vs[j * dp + c], acc);
            op[c] = from_f32<T>(acc * inv);
        }
        __syncwarp();
    }
}

// Backward. grid = (heads, nchunks); each CTA walks windows chunk, chunk + nchunks, ... of its head and keeps the
// bias gradient of those windows in shared memory, flushing it with one atomicAdd per element at the end.
template <typename T>
__global__ void __launch_bounds__(NW * 32) swin_attn_bwd_kernel(SwinArgs a) {
    extern __shared__ float sm[];
    const int N = a.w * a.w, d = a.d, dp = d + 1;
    float* qs = sm;                                   // normalised q
    float* ks = qs + N * dp;                          // normalised k
    float* vs = ks + N * dp;
    float* dos = vs + N * dp;
    float* dbias_s = dos + N * dp;                    // [N][N]
    float* pbuf = dbias_s + N * N;                    // [NW][N]
    float* dsbuf = pbuf + NW * N;                     // [NW][N]
    float* qn = dsbuf + NW * N;                       // [N] max(|q|, eps)
    float* kn = qn + N;                               // [N]
    float* lse_s = kn + N;                            // [N]
    float* d_s = lse_s + N;                           // [N]
    int* tok = reinterpret_cast<int*>(d_s + N);       // [N]
    int* reg = tok + N;                               // [N]
    __shared__ float s_dscale;
    const int nW = (a.res / a.w) * (a.res / a.w);
    const int h = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float raw_ls = a.logit_scale[h];
    const float scale = __expf(fminf(raw_ls, LOGIT_MAX));
    const float* bias = a.bias + static_cast<long long>(h) * N * N;
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) dbias_s[idx] = 0.0f;
    if (threadIdx.x == 0) s_dscale = 0.0f;
    float dscale_acc = 0.0f;
    float* pw = pbuf + warp * N;
    float* dsw = dsbuf + warp * N;

    for (int bw = blockIdx.y; bw < a.B * nW; bw += a.nchunks) {
        const int b = bw / nW, win = bw % nW;
        __syncthreads();
        for (int n = threadIdx.x; n < N; n += blockDim.x) window_token(a, b, win, n, tok[n], reg[n]);
        __syncthreads();
        for (int n = warp; n < N; n += NW) {
            const long long base = static_cast<long long>(tok[n]) * a.ld + h * d;
            const long long cbase = static_cast<long long>(tok[n]) * a.ldc + h * d;
            float sq = 0.0f, sk = 0.0f, dd = 0.0f;
            for (int c = lane; c < d; c += 32) {
                const float qv = to_f32(reinterpret_cast<const T*>(a.q)[base + c]);
                const float kv = to_f32(reinterpret_cast<const T*>(a.k)[base + c]);
                const float g = to_f32(reinterpret_cast<const T*>(a.dctx)[cbase + c]);
                qs[n * dp + c] = qv;
                ks[n * dp + c] = kv;
                vs[n * dp + c] = to_f32(reinterpret_cast<const T*>(a.v)[base + c]);
                dos[n * dp + c] = g;
                dd += g * to_f32(reinterpret_cast<const T*>(a.ctx)[cbase + c]);
                sq += qv * qv;
                sk += kv * kv;
            }
            const float nq = fmaxf(sqrtf(warp_sum(sq)), NORM_EPS);
            const float nk = fmaxf(sqrtf(warp_sum(sk)), NORM_EPS);
            dd = warp_sum(dd);
            for (int c = lane; c < d; c += 32) {
                qs[n * dp + c] /= nq;
                ks[n * dp + c] /= nk;
            }
            if (lane == 0) {
                qn[n] = nq;
                kn[n] = nk;
                d_s[n] = dd;
                lse_s[n] = a.lse[(static_cast<long long>(bw) * a.heads + h) * N + n];
            }
        }
        __syncthreads();
        // ---- pass A: per query row -> dq, dbias, dscale ----
        for (int i = warp; i < N; i += NW) {
            for (int j = lane; j < N; j += 32) {
                float cs = 0.0f, dpv = 0.0f;
                for (int c = 0; c < d; ++c) {
                    cs = fmaf(qs[i * dp + c], ks[j * dp + c], cs);
                    dpv = fmaf(dos[i * dp + c], vs[j * dp + c], dpv);
                }
                float s = cs * scale + bias[i * N + j];
                if (reg[i] != reg[j]) s += -200.0f;
                const float p = __expf(s - lse_s[i]);
                const float ds = p * (dpv - d_s[i]);
                dsw[j] = ds;
                dbias_s[i * N + j] += ds;          // row i of this window is owned by this warp: no race
                dscale_acc += ds * cs;
            }
            __syncwarp();
            // dqhat = scale * sum_j ds_j khat_j ; dq = (dqhat - qhat (qhat . dqhat)) / |q|
            float dot = 0.0f;
            float dqh[4];                           // d <= 128 -> up to 4 channels per lane
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int c = lane + 32 * t;
                float acc = 0.0f;
                if (c < d) {
                    for (int j = 0; j < N; ++j) acc = fmaf(dsw[j], ks[j * dp + c], acc);
                    acc *= scale;
                    dot += acc * qs[i * dp + c];
                }
                dqh[t] = acc;
            }
            dot = warp_sum(dot);
            T* dqp = reinterpret_cast<T*>(a.dq) + static_cast<long long>(tok[i]) * a.ld + h * d;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int c = lane + 32 * t;
                if (c < d) dqp[c] = from_f32<T>((dqh[t] - qs[i * dp + c] * dot) / qn[i]);
            }
            __syncwarp();
        }
        // ---- pass B: per key row -> dk, dv ----
        for (int j = warp; j < N; j += NW) {
            for (int i = lane; i < N; i += 32) {
                float cs = 0.0f, dpv = 0.0f;
                for (int c = 0; c < d; ++c) {
                    cs = fmaf(qs[i * dp + c], ks[j * dp + c], cs);
                    dpv = fmaf(dos[i * dp + c], vs[j * dp + c], dpv);
                }
                float s = cs * scale + bias[i * N + j];
                if (reg[i] != reg[j]) s += -200.0f;
                const float p = __expf(s - lse_s[i]);
                pw[i] = p;
                dsw[i] = p * (dpv - d_s[i]);
            }
            __syncwarp();
            float dot = 0.0f;
            float dkh[4], dvv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int c = lane + 32 * t;
                float acck = 0.0f, accv = 0.0f;
                if (c < d) {
                    for (int i = 0; i < N; ++i) {
                        acck = fmaf(dsw[i], qs[i * dp + c], acck);
                        accv = fmaf(pw[i], dos[i * dp + c], accv);
                    }
                    acck *= scale;
                    dot += acck * ks[j * dp + c];
                }
                dkh[t] = acck;
                dvv[t] = accv;
            }
            dot = warp_sum(dot);
            const long long base = static_cast<long long>(tok[j]) * a.ld + h * d;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int c = lane + 32 * t;
                if (c < d) {
                    reinterpret_cast<T*>(a.dk)[base + c] = from_f32<T>((dkh[t] - ks[j * dp + c] * dot) / kn[j]);
                    reinterpret_cast<T*>(a.dv)[base + c] = from_f32<T>(dvv[t]);
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    float* dbias = a.dbias + static_cast<long long>(h) * N * N;
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) atomicAdd(&dbias[idx], dbias_s[idx]);
    dscale_acc = warp_sum(dscale_acc);
    if (lane == 0) atomicAdd(&s_dscale, dscale_acc);
    __syncthreads();
    // d(logit_scale) = dscale * exp(clamped) * [logit_scale < ln 100]   (torch.clamp passes gradient at equality)
    if (threadIdx.x == 0) atomicAdd(&a.dlogit_scale[h], raw_ls <= LOGIT_MAX ? s_dscale * scale : 0.0f);
}

// ---------------------------------------------------------------------------------------------------------
// Continuous position bias (:408-410, :450-460).  T = (2w-1)^2 table rows, U = 512 hidden units.
//   hidden[t,u] = relu(W1[u,:] . coords[t,:] + b1[u]);  tab[t,h] = W2[h,:] . hidden[t,:]
//   bias[h,i,j] = 16 * sigmoid(tab[index[i,j], h])
// ---------------------------------------------------------------------------------------------------------
__global__ void cpb_fwd_kernel(const float* __restrict__ coords, const float* __restrict__ w1, const float* __restrict__ b1,
                               const float* __restrict__ w2, int U, int heads, float* __restrict__ hidden, float* __restrict__ tab) {
    extern __shared__ float hs[];                       // [U]
    const int t = blockIdx.x;
    const float c0 = coords[2 * t], c1 = coords[2 * t + 1];
    for (int u = threadIdx.x; u < U; u += blockDim.x) {
        const float v = fmaxf(fmaf(w1[2 * u], c0, fmaf(w1[2 * u + 1], c1, b1[u])), 0.0f);
        hs[u] = v;
        hidden[static_cast<long long>(t) * U + u] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int h = warp; h < heads; h += nwarps) {
        float s = 0.0f;
        for (int u = lane; u < U; u += 32) s = fmaf(hs[u], w2[static_cast<long long>(h) * U + u], s);
        s = warp_sum(s);
        if (lane == 0) tab[t * heads + h] = s;
    }
}

__global__ void cpb_gather_kernel(const float* __restrict__ tab, const int* __restrict__ index, int NN, int heads, float* __restrict__ bias) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= NN * heads) return;
    const int h = idx / NN, ij = idx % NN;
    const float x = tab[index[ij] * heads + h];
    bias[idx] = 16.0f / (1.0f + __expf(-x));
}

// dtab[t,h] = sum_{(i,j): index = t} dbias[h,i,j] * 16 s (1 - s)
__global__ void cpb_bwd_scatter_kernel(const float* __restrict__ dbias, const float* __restrict__ tab, const int* __restrict__ index,
                                       int NN, int heads, float* __restrict__ dtab) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= NN * heads) return;
    const int h = idx / NN, ij = idx % NN;
    const int t = index[ij];
    const float s = 1.0f / (1.0f + __expf(-tab[t * heads + h]));
    atomicAdd(&dtab[t * heads + h], dbias[idx] * 16.0f * s * (1.0f - s));
}

// dW2[h,u] = sum_t dtab[t,h] hidden[t,u]
__global__ void cpb_bwd_w2_kernel(const float* __restrict__ dtab, const float* __restrict__ hidden, int Tn, int U, int heads,
                                  float* __restrict__ dw2, int accumulate) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= heads * U) return;
    const int h = idx / U, u = idx % U;
    float s = 0.0f;
    for (int t = 0; t < Tn; ++t) s = fmaf(dtab[t * heads + h], hidden[static_cast<long long>(t) * U + u], s);
    dw2[idx] = accumulate ? dw2[idx] + s : s;
}

// dhidden[t,u] = [hidden > 0] sum_h dtab[t,h] W2[h,u];  dW1[u,:] = sum_t dhidden[t,u] coords[t,:];  db1[u] = sum_t dhidden[t,u]
// one CTA per hidden unit u, threads over table rows t
__global__ void __launch_bounds__(128)
cpb_bwd_w1_kernel(const float* __restrict__ dtab, const float* __restrict__ hidden, const float* __restrict__ w2,
                  const float* __restrict__ coords, int Tn, int U, int heads, float* __restrict__ dw1,
                  float* __restrict__ db1, int accumulate) {
    __shared__ float red[3][4];
    const int u = blockIdx.x;
    float g0 = 0.0f, g1 = 0.0f, gb = 0.0f;
    for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
        if (hidden[static_cast<long long>(t) * U + u] > 0.0f) {
            float dh = 0.0f;
            for (int h = 0; h < heads; ++h) dh = fmaf(dtab[t * heads + h], w2[static_cast<long long>(h) * U + u], dh);
            g0 = fmaf(dh, coords[2 * t], g0);
            g1 = fmaf(dh, coords[2 * t + 1], g1);
            gb += dh;
        }
    }
    g0 = warp_sum(g0); g1 = warp_sum(g1); gb = warp_sum(gb);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = g0; red[1][threadIdx.x >> 5] = g1; red[2][threadIdx.x >> 5] = gb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        g0 = red[0][0] + red[0][1] + red[0][2] + red[0][3];
        g1 = red[1][0] + red[1][1] + red[1][2] + red[1][3];
        gb = red[2][0] + red[2][1] + red[2][2] + red[2][3];
        if (accumulate) {
            dw1[2 * u] += g0; dw1[2 * u + 1] += g1; db1[u] += gb;
        } else {
            dw1[2 * u] = g0; dw1[2 * u + 1] = g1; db1[u] = gb;
        }
    }
}

size_t swin_fwd_smem(int N, int d) { return sizeof(float) * (3ll * N * (d + 1) + NW * N) + sizeof(int) * 2ll * N; }
size_t swin_bwd_smem(int N, int d) {
    return sizeof(float) * (4ll * N * (d + 1) + 1ll * N * N + 2ll * NW * N + 4ll * N) + sizeof(int) * 2ll * N;
}

}  // namespace

// tensor-core path (swin_attention_tc.cu)
bool swin_attention_tc_supported(int dtype, int head_dim, int window, long long ld, long long ldc, const void* q, const void* k,
                                 const void* v, const void* ctx);
int swin_attention_fwd_tc(cudaStream_t st, int B, int res, int heads, int window, int shift, const void* q, const void* k, const void* v,
                          long long ld, void* ctx, long long ldc, const float* logit_scale, const float* bias, float* lse);
int swin_attention_bwd_tc(cudaStream_t st, int B, int res, int heads, int window, int shift, const void* q, const void* k, const void* v,
                          long long ld, const void* ctx, const void* dctx, long long ldc, void* dq, void* dk, void* dv,
                          const float* logit_scale, const float* bias, const float* lse, float* dbias, float* dlogit_scale);
// tensor-core path for windows of 65..144 tokens (swin_attention_tc_big.cu)
bool swin_attention_big_supported(int dtype, int head_dim, int window, long long ld, long long ldc, const void* q, const void* k,
                                  const void* v, const void* ctx);
int swin_attention_fwd_big(cudaStream_t st, int B, int res, int heads, int window, int shift, const void* q, const void* k, const void* v,
                           long long ld, void* ctx, long long ldc, const float* logit_scale, const float* bias, float* lse);
int swin_attention_bwd_big(cudaStream_t st, int B, int res, int heads, int window, int shift, const void* q, const void* k, const void* v,
                           long long ld, const void* ctx, const void* dctx, long long ldc, void* dq, void* dk, void* dv,
                           const float* logit_scale, const float* bias, const float* lse, float* dbias, float* dlogit_scale);
}  // namespace klab

using namespace klab;

static bool swin_force_generic() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("KLAB_ATTENTION_GENERIC");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

extern "C" {

int klab_swin_attention_fwd(void* stream, int dtype, int B, int res, int heads, int head_dim, int window, int shift,
                            const void* q, const void* k, const void* v, long long ld, void* ctx, long long ldc,
                            const float* logit_scale, const float* bias, float* lse) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(B > 0 && res > 0 && window > 0 && res % window == 0, "swin_attention_fwd: grid %d is not a multiple of window %d", res, window);
    KLAB_REQUIRE(head_dim <= 128, "swin_attention: head_dim %d > 128", head_dim);
    if (!swin_force_generic() && swin_attention_tc_supported(dtype, head_dim, window, ld, ldc, q, k, v, ctx))
        return swin_attention_fwd_tc(static_cast<cudaStream_t>(stream), B, res, heads, window, shift, q, k, v, ld, ctx, ldc, logit_scale, bias, lse);
    if (!swin_force_generic() && swin_attention_big_supported(dtype, head_dim, window, ld, ldc, q, k, v, ctx))
        return swin_attention_fwd_big(static_cast<cudaStream_t>(stream), B, res, heads, window, shift, q, k, v, ld, ctx, ldc, logit_scale, bias, lse);
    SwinArgs a{};
    a.q = q; a.k = k; a.v = v; a.out = ctx; a.ld = ld; a.ldc = ldc;
    a.B = B; a.res = res; a.heads = heads; a.d = head_dim; a.w = window; a.shift = shift;
    a.logit_scale = logit_scale; a.bias = bias; a.lse = lse;
    const int N = window * window, nW = (res / window) * (res / window);
    const size_t smem = swin_fwd_smem(N, head_dim);
    KLAB_REQUIRE(smem <= 227 * 1024, "swin_attention_fwd: window too large (%zu bytes smem)", smem);
    const dim3 grid(B * nW, heads);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == KLAB_BF16) {
        KLAB_CHECK_CUDA(cudaFuncSetAttribute(swin_attn_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        swin_attn_fwd_kernel<__nv_bfloat16><<<grid, NW * 32, smem, st>>>(a);
    } else {
        KLAB_CHECK_CUDA(cudaFuncSetAttribute(swin_attn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        swin_attn_fwd_kernel<float><<<grid, NW * 32, smem, st>>>(a);
    }
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

// dbias [heads,N,N] and dlogit_scale [heads] are OVERWRITTEN (zeroed here, then accumulated by the kernel).
int klab_swin_attention_bwd(void* stream, int dtype, int B, int res, int heads, int head_dim, int window, int shift,
                            const void* q, const void* k, const void* v, long long ld, const void* ctx, const void* dctx,
                            long long ldc, void* dq, void* dk, void* dv, const float* logit_scale, const float* bias,
                            const float* lse, float* dbias, float* dlogit_scale) {
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(B > 0 && res > 0 && window > 0 && res % window == 0, "swin_attention_bwd: grid %d is not a multiple of window %d", res, window);
    KLAB_REQUIRE(head_dim <= 128, "swin_attention: head_dim %d > 128", head_dim);
    if (!swin_force_generic() && swin_attention_tc_supported(dtype, head_dim, window, ld, ldc, q, k, v, ctx) &&
        (reinterpret_cast<uintptr_t>(dctx) & 15) == 0)
        return swin_attention_bwd_tc(static_cast<cudaStream_t>(stream), B, res, heads, window, shift, q, k, v, ld, ctx, dctx, ldc, dq, dk, dv,
                                     logit_scale, bias, lse, dbias, dlogit_scale);
    if (!swin_force_generic() && swin_attention_big_supported(dtype, head_dim, window, ld, ldc, q, k, v, ctx) &&
        (reinterpret_cast<uintptr_t>(dctx) & 15) == 0)
        return swin_attention_bwd_big(static_cast<cudaStream_t>(stream), B, res, heads, window, shift, q, k, v, ld, ctx, dctx, ldc, dq, dk, dv,
                                      logit_scale, bias, lse, dbias, dlogit_scale);
    SwinArgs a{};
    a.q = q; a.k = k; a.v = v; a.ctx = ctx; a.dctx = dctx; a.dq = dq; a.dk = dk; a.dv = dv; a.ld = ld; a.ldc = ldc;
    a.B = B; a.res = res; a.heads = heads; a.d = head_dim; a.w = window; a.shift = shift;
    a.logit_scale = logit_scale; a.bias = bias; a.lse = const_cast<float*>(lse);
    a.dbias = dbias; a.dlogit_scale = dlogit_scale;
    const int N = window * window, nW = (res / window) * (res / window);
    int nchunks = (4 * sm_count() + heads - 1) / heads;
    if (nchunks > B * nW) nchunks = B * nW;
    a.nchunks = nchunks;
    const size_t smem = swin_bwd_smem(N, head_dim);
    KLAB_REQUIRE(smem <= 227 * 1024, "swin_attention_bwd: window too large (%zu bytes smem)", smem);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    KLAB_CHECK_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * heads * N * N, st));
    KLAB_CHECK_CUDA(cudaMemsetAsync(dlogit_scale, 0, sizeof(float) * heads, st));
    const dim3 grid(heads, nchunks);
    if (dtype == KLAB_BF16) {
        KLAB_CHECK_CUDA(cudaFuncSetAttribute(swin_attn_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        swin_attn_bwd_kernel<__nv_bfloat16><<<grid, NW * 32, smem, st>>>(a);
    } else {
        KLAB_CHECK_CUDA(cudaFuncSetAttribute(swin_attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        swin_attn_bwd_kernel<float><<<grid, NW * 32, smem, st>>>(a);
    }
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

// bias[h,i,j] = 16 sigmoid(MLP(coords)[index[i,j], h]); hidden [(2w-1)^2, U] and tab [(2w-1)^2, heads] are kept for backward.
int klab_swin_cpb_fwd(void* stream, int table_rows, int hidden_units, int heads, int n_tokens, const float* coords,
                      const int* index, const float* w1, const float* b1, const float* w2, float* hidden, float* tab, float* bias) {
    if (int rc = klab_check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cpb_fwd_kernel<<<table_rows, 256, sizeof(float) * hidden_units, st>>>(coords, w1, b1, w2, hidden_units, heads, hidden, tab);
    KLAB_LAUNCH_CHECK();
    const int NN = n_tokens * n_tokens, total = NN * heads;
    cpb_gather_kernel<<<(total + 255) / 256, 256, 0, st>>>(tab, index, NN, heads, bias);
    KLAB_LAUNCH_CHECK();
    count_launch(2);
    return KLAB_OK;
}

// dtab is scratch [(2w-1)^2, heads]
int klab_swin_cpb_bwd(void* stream, int table_rows, int hidden_units, int heads, int n_tokens, const float* coords,
                      const int* index, const float* w2, const float* hidden, const float* tab, const float* dbias, float* dtab,
                      float* dw1, float* db1, float* dw2, int accumulate) {
    if (int rc = klab_check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    KLAB_CHECK_CUDA(cudaMemsetAsync(dtab, 0, sizeof(float) * table_rows * heads, st));
    const int NN = n_tokens * n_tokens, total = NN * heads;
    cpb_bwd_scatter_kernel<<<(total + 255) / 256, 256, 0, st>>>(dbias, tab, index, NN, heads, dtab);
    KLAB_LAUNCH_CHECK();
    cpb_bwd_w2_kernel<<<(heads * hidden_units + 127) / 128, 128, 0, st>>>(dtab, hidden, table_rows, hidden_units, heads, dw2, accumulate);
    KLAB_LAUNCH_CHECK();
    cpb_bwd_w1_kernel<<<hidden_units, 128, 0, st>>>(dtab, hidden, w2, coords, table_rows, hidden_units, heads, dw1, db1, accumulate);
    KLAB_LAUNCH_CHECK();
    count_launch(3);
    return KLAB_OK;
}

}  // extern "C"
