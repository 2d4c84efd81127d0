// Helpers for hand-staged UMMA operand tiles: [rows x 64] bf16 tiles in the 128-byte swizzle that TMA SWIZZLE_128B and the
// UMMA SWIZZLE_128B descriptors share (16-byte chunk index XOR (row % 8); tiles are 1024-byte aligned).
#pragma once

#include "common.cuh"

namespace klab {

__device__ __forceinline__ uint32_t sw128(int row, int col) {
    return static_cast<uint32_t>(row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1)));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// store 8 consecutive columns [col8*8, col8*8+8) of row `row`
__device__ __forceinline__ void st_tile8(uint8_t* tile, int row, int col8, const float* v) {
    uint4 q;
    q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]); q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(tile + row * 128 + ((col8 ^ (row & 7)) << 4)) = q;
}
__device__ __forceinline__ void st_tile8_raw(uint8_t* tile, int row, int col8, uint4 q) {
    *reinterpret_cast<uint4*>(tile + row * 128 + ((col8 ^ (row & 7)) << 4)) = q;
}
__device__ __forceinline__ uint4 ld_tile8_raw(const uint8_t* tile, int row, int col8) {
    return *reinterpret_cast<const uint4*>(tile + row * 128 + ((col8 ^ (row & 7)) << 4));
}

__device__ __forceinline__ void unpack8(uint4 q, float* v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float2 f = __bfloat1622float2(h[t]);
        v[2 * t] = f.x;
        v[2 * t + 1] = f.y;
    }
}

__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 q;
    q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]); q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    return q;
}

}  // namespace klab
