// Shared pieces of the tensor-core Swin-V2 window-attention kernels (swin_attention_tc.cu: windows of <= 64 tokens, two per
// 128-row tile; swin_attention_tc_big.cu: windows of 65..192 tokens, e.g. the 12 x 12 windows of 384^2 inputs).
#pragma once

#include "tc_tiles.cuh"

namespace klab {
namespace swintc {

constexpr int HD = 32;
constexpr int TILE = 128;
constexpr int SLOT = 64;
constexpr int TB = TILE * 128;                 // bytes of one [128 x 64] bf16 tile
constexpr float LOGIT_MAX = 4.605170185988092f;
constexpr float NORM_EPS = 1e-12f;

struct SwinTcArgs {
    const __nv_bfloat16 *q, *k, *v, *ctx, *dctx;
    __nv_bfloat16 *out, *dq, *dk, *dv;
    long long ld, ldc;
    int B, res, heads, w, shift, N, nW;
    const float* logit_scale;
    const float* bias;
    float* lse;
    float* dbias;
    float* dlogit_scale;
};

__device__ __forceinline__ int region_of(int y, int res, int w, int shift) { return y < res - w ? 0 : (y < res - shift ? 1 : 2); }

// token row index and shift-mask region of token n of global window bw (-1 if the slot is padding)
__device__ __forceinline__ int window_token(const SwinTcArgs& a, int bw, int n, int& region) {
    region = 0;
    if (n >= a.N || bw >= a.B * a.nW) return -1;
    const int b = bw / a.nW, win = bw % a.nW;
    const int nwx = a.res / a.w;
    const int ys = (win / nwx) * a.w + n / a.w, xs = (win % nwx) * a.w + n % a.w;
    int y = ys + a.shift, x = xs + a.shift;
    if (y >= a.res) y -= a.res;
    if (x >= a.res) x -= a.res;
    if (a.shift > 0) region = region_of(ys, a.res, a.w, a.shift) * 3 + region_of(xs, a.res, a.w, a.shift);
    return (b * a.res + y) * a.res + x;
}

// load one 32-element bf16 head slice (64 B) of a token row
__device__ __forceinline__ void load_row32(const __nv_bfloat16* base, long long ld, int tok, int h, float* v) {
    const uint4* p = reinterpret_cast<const uint4*>(base + static_cast<long long>(tok) * ld + h * HD);
#pragma unroll
    for (int c = 0; c < 4; ++c) unpack8(p[c], v + 8 * c);
}
__device__ __forceinline__ void store_row32(__nv_bfloat16* base, long long ld, int tok, int h, const float* v) {
    uint4* p = reinterpret_cast<uint4*>(base + static_cast<long long>(tok) * ld + h * HD);
#pragma unroll
    for (int c = 0; c < 4; ++c) p[c] = pack8(v + 8 * c);
}
__device__ __forceinline__ void stage_row32(uint8_t* tile, int row, const float* v) {
#pragma unroll
    for (int c = 0; c < 4; ++c) st_tile8(tile, row, c, v + 8 * c);
}
// Stage a unit vector as bf16 hi (columns 0..31) + bf16 lo = v - hi (columns 32..63).  The cosine logits are multiplied by
// up to 100 (logit_scale clamp) before the softmax, so S = Qh Kh^T is accumulated as hi*hi + lo*hi + hi*lo (three
// K = 32 tcgen05 passes): ~16 mantissa bits on the operands instead of 8, at negligible cost for a K = 32 product.
__device__ __forceinline__ void stage_row32_hilo(uint8_t* tile, int row, const float* v) {
    float lo[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) lo[c] = v[c] - __bfloat162float(__float2bfloat16_rn(v[c]));
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        st_tile8(tile, row, c, v + 8 * c);
        st_tile8(tile, row, 4 + c, lo + 8 * c);
    }
}
// S[tmem] = Qhi Khi^T + Qlo Khi^T + Qhi Klo^T  (qa / ka: shared addresses of the hi|lo tiles)
__device__ __forceinline__ void issue_cosine_logits(uint32_t tmem_s, uint32_t qa, uint32_t ka, uint32_t idesc) {
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t qo = pass == 1 ? 64 : 0, ko = pass == 2 ? 64 : 0;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
            umma_bf16(tmem_s, umma_smem_desc_sw128(qa + qo + k * 32, 16, 1024), umma_smem_desc_sw128(ka + ko + k * 32, 16, 1024), idesc,
                      (pass | k) != 0);
    }
}

__device__ __forceinline__ float inv_norm32(const float* v, float& norm) {
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < HD; ++c) s = fmaf(v[c], v[c], s);
    norm = fmaxf(sqrtf(s), NORM_EPS);
    return 1.0f / norm;
}

}  // namespace swintc
}  // namespace klab
