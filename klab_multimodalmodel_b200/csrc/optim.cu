// N1: the optimizer step that follows the hot path, /root/reference/train.py:28,66 -- torch.optim.Adam over
// model.transformer.parameters().  One launch updates EVERY parameter tensor (multi-tensor apply): the step is pure HBM
// traffic (read p, g, m, v; write p, m, v = 28 B per parameter, 20.7 GB for T5-large), so the only things that matter are
// 16-byte accesses, enough bytes in flight per SM and not paying 250 launches.  Optionally the kernel also refreshes the
// bf16 operand copy of the parameter (what the tensor cores read next step), which removes the separate cast pass.
//
// Semantics are those of torch.optim.Adam (torch/optim/adam.py, _single_tensor_adam; amsgrad = False, maximize = False):
//   g' = g + wd * p;  m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace klab {
void count_launch(int n = 1);
namespace {

constexpr int ADAM_CHUNK = 16384;     // elements per CTA

struct AdamArgs {
    float lr, beta1, beta2, eps, weight_decay;
    float omb1, omb2;     // 1 - beta, rounded from the DOUBLE difference (1.0f - 0.98f is off by 1e-6 relative)
    float inv_bc1;        // 1 / (1 - beta1^t)
    float inv_sqrt_bc2;   // 1 / sqrt(1 - beta2^t)
    float grad_scale;     // multiplies g first (1 = off): gradient-accumulation averaging without a separate pass
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a) {
    g *= a.grad_scale;
    g = fmaf(a.weight_decay, p, g);
    m = fmaf(a.beta1, m, a.omb1 * g);
    v = fmaf(a.beta2, v, a.omb2 * g * g);
    const float denom = fmaf(sqrtf(v), a.inv_sqrt_bc2, a.eps);
    p -= (a.lr * a.inv_bc1) * (m / denom);
}

// table[t] = {p, g, m, v, bf16 copy (or 0), numel}; blockmap[b] = {tensor index, chunk index}
// (<= 40 registers per thread, small blocks: a block then fits into what resident CTAs of the step's big kernels leave free on an SM
//  -- 10 240 registers next to a tcgen05 GEMM CTA (576 threads x 96), ~7 000 next to two Swin attention CTAs -- so the update, issued
//  on a side stream by optim.py, can make progress WHILE the next step's kernels run, not only in the gaps between them)
template <int ADAM_THREADS>
__global__ void __launch_bounds__(ADAM_THREADS, 1536 / ADAM_THREADS) adam_multi_kernel(const long long* __restrict__ table, const int* __restrict__ blockmap,
                                                                                      AdamArgs a) {
    const int t = blockmap[2 * blockIdx.x], chunk = blockmap[2 * blockIdx.x + 1];
    const long long* e = table + 6ll * t;
    float* __restrict__ p = reinterpret_cast<float*>(e[0]);
    const float* __restrict__ g = reinterpret_cast<const float*>(e[1]);
    float* __restrict__ m = reinterpret_cast<float*>(e[2]);
    float* __restrict__ v = reinterpret_cast<float*>(e[3]);
    __nv_bfloat16* __restrict__ w16 = reinterpret_cast<__nv_bfloat16*>(e[4]);
    const long long n = e[5];
    const long long lo = static_cast<long long>(chunk) * ADAM_CHUNK;
    const long long hi = lo + ADAM_CHUNK < n ? lo + ADAM_CHUNK : n;
    const bool vec = ((e[0] | e[1] | e[2] | e[3]) & 15) == 0 && (e[4] & 7) == 0;
    if (vec) {
        const long long hi4 = lo + ((hi - lo) & ~3ll);
        for (long long i = lo + 4ll * threadIdx.x; i < hi4; i += 4ll * ADAM_THREADS) {
            float4 P = *reinterpret_cast<float4*>(p + i);
            const float4 G = __ldcs(reinterpret_cast<const float4*>(g + i));          // gradients are dead after this read
            float4 M = *reinterpret_cast<float4*>(m + i);
            float4 V = *reinterpret_cast<float4*>(v + i);
            adam_one(P.x, G.x, M.x, V.x, a);
            adam_one(P.y, G.y, M.y, V.y, a);
            adam_one(P.z, G.z, M.z, V.z, a);
            adam_one(P.w, G.w, M.w, V.w, a);
            *reinterpret_cast<float4*>(p + i) = P;
            *reinterpret_cast<float4*>(m + i) = M;
            *reinterpret_cast<float4*>(v + i) = V;
            if (w16) {
                const __nv_bfloat162 h0 = __floats2bfloat162_rn(P.x, P.y), h1 = __floats2bfloat162_rn(P.z, P.w);
                uint2 q;
                q.x = *reinterpret_cast<const uint32_t*>(&h0);
                q.y = *reinterpret_cast<const uint32_t*>(&h1);
                *reinterpret_cast<uint2*>(w16 + i) = q;
            }
        }
        for (long long i = hi4 + threadIdx.x; i < hi; i += ADAM_THREADS) {
            float P = p[i], M = m[i], V = v[i];
            adam_one(P, g[i], M, V, a);
            p[i] = P; m[i] = M; v[i] = V;
            if (w16) w16[i] = __float2bfloat16_rn(P);
        }
    } else {
        for (long long i = lo + threadIdx.x; i < hi; i += ADAM_THREADS) {
            float P = p[i], M = m[i], V = v[i];
            adam_one(P, g[i], M, V, a);
            p[i] = P; m[i] = M; v[i] = V;
            if (w16) w16[i] = __float2bfloat16_rn(P);
        }
    }
}

}  // namespace
}  // namespace klab

extern "C" {

int klab_adam_chunk_elems(void) { return klab::ADAM_CHUNK; }

int klab_adam_step(void* stream, const long long* table_dev, const int* blockmap_dev, int n_blocks, double lr, double beta1,
                   double beta2, double eps, double weight_decay, long long step, double grad_scale) {
    using namespace klab;
    if (int rc = klab_check_device()) return rc;
    KLAB_REQUIRE(n_blocks > 0 && step >= 1 && table_dev && blockmap_dev, "adam_step: bad arguments (n_blocks=%d step=%lld)", n_blocks, step);
    AdamArgs a;
    a.beta1 = static_cast<float>(beta1); a.beta2 = static_cast<float>(beta2); a.eps = static_cast<float>(eps);
    a.weight_decay = static_cast<float>(weight_decay); a.grad_scale = static_cast<float>(grad_scale);
    a.omb1 = static_cast<float>(1.0 - beta1); a.omb2 = static_cast<float>(1.0 - beta2);
    a.lr = static_cast<float>(lr / (1.0 - pow(beta1, static_cast<double>(step))));      // step size lr / (1 - b1^t), formed in double
    a.inv_bc1 = 1.0f;
    a.inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(1.0 - pow(beta2, static_cast<double>(step))));
    static const int threads = []() { const char* e = getenv("KLAB_ADAM_THREADS"); return e && atoi(e) == 256 ? 256 : 128; }();
    if (threads == 256) adam_multi_kernel<256><<<n_blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(table_dev, blockmap_dev, a);
    else adam_multi_kernel<128><<<n_blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(table_dev, blockmap_dev, a);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

}  // extern "C"
