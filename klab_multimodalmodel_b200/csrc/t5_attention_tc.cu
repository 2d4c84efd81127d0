// K9 (hot path): T5 attention on the 5th-gen tensor cores, d_kv = 64, bf16 operands, whole (batch, head) problem resident.
//
// Semantics are exactly those of t5_attention.cu (HF/models/t5/modeling_t5.py:253-344: unscaled q k^T + bucketed relative bias
// [+ causal mask], fp32 softmax, dropout on the probabilities, P v); this file only changes WHERE the arithmetic runs:
//   * Q / K / V (and dO) tiles are fetched by TMA straight from the [B*L, H*64] projection-GEMM layout into 128B-swizzled
//     shared memory -- no head transpose, no padding copies (rows past the sequence end are masked, not copied);
//   * S = Q K^T, dP = dO V^T, O = P V, dV = P^T dO, dK = dS^T Q, dQ = dS K are tcgen05.mma with fp32 accumulators in TMEM;
//     one smem tile serves both as a K-major and as an MN-major operand (P as A of P V and as A^T of P^T dO, ...);
//   * softmax / dS run one-thread-per-row out of TMEM (tcgen05.ld), so S, P and dS never touch HBM.
// Limits: d_kv == 64, Lq <= 256, Lk <= 256 (the reference's sequences are <= 176); anything else takes the generic kernel.
#include <cstdlib>

#include "common.cuh"

namespace klab {
void count_launch(int n = 1);
int sm_count();
namespace {

constexpr int DK = 64;
constexpr int TILE = 128;

struct TcArgs {
    void* out;                 // fwd: ctx [B*Lq, ldo]
    const void* o;             // bwd: ctx
    const void* dout;          // bwd
    void *dq, *dk, *dv;        // bwd outputs (row strides ldq / ldk / ldv)
    long long ldq, ldk, ldv, ldo;
    int B, H, Lq, Lk;
    const float* bias_table;   // [num_buckets, H] or null
    const int* rel_bucket;     // LUT, index (j - i - q_offset) + rel_zero
    int rel_zero, num_buckets, causal, q_offset;
    float* lse;                // [B, H, Lq]
    float* dbias_partial;      // bwd: [B*H, num_buckets]
    float dropout_p;
    unsigned long long seed;
    const unsigned long long* seed_ptr;
    int* sched;                // dynamic work distribution of the persistent kernels ({next problem, finished CTAs}) or null
    int pack;                  // single-tile kernels: G = 128 / L problems (consecutive batch entries of one head) share one tile; 1 = off
};

// byte offset of element (row, col) inside a [rows x 64] bf16 tile stored with the 128-byte swizzle (TMA SWIZZLE_128B /
// UMMA SWIZZLE_128B): 16-byte chunk index XOR (row % 8)
__device__ __forceinline__ uint32_t sw128(int row, int col) {
    return static_cast<uint32_t>(row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1)));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// store 8 consecutive columns [col8*8, col8*8+8) of row `row` of a swizzled [128 x 64] tile
__device__ __forceinline__ void st_tile8(uint8_t* tile, int row, int col8, const float* v) {
    uint4 q;
    q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]); q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(tile + row * 128 + ((col8 ^ (row & 7)) << 4)) = q;
}

__device__ __forceinline__ float drop_mult(const TcArgs& a, const DropKey& key, int bh, int i, int j) {
    const uint64_t idx = (static_cast<uint64_t>(bh) * a.Lq + i) * a.Lk + j;
    return dropout_mult(key, idx);
}

// Bias gradient of the single-tile backward kernels: dBias[bucket(j - i)] += dS_ij, i.e. sums of the dS tile along its diagonals.
// A warp holds a 32 x 16 block of dS in registers (lane = row, v[t] = column t of the chunk): lane m collects diagonal m (elements with
// row <= column) and diagonal m - 32 (row > column) -- at step t it takes column t from lane (t - m) mod 32, every element is taken by
// exactly one lane -- and adds the two sums to drel[], indexed by tile column - tile row + 127.  16 shuffles per chunk, no shared-memory
// reads.  (The sums used to be formed after the fact by one thread per diagonal walking the swizzled bf16 dS tile: 30 % of the
// kernel's executed instructions, and the longest diagonals set the critical path -- profiles/r02_t5_attn_final_ncu.txt.)
__device__ __forceinline__ void diag_accumulate16(float* drel, const float (&v)[16], int dd0, int lane) {
    float acc_p = 0.0f, acc_n = 0.0f;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const float x = __shfl_sync(0xffffffffu, v[t], (t - lane) & 31);
        if (lane <= t) acc_p += x;
        else acc_n += x;
    }
    if (lane < 16) atomicAdd(&drel[dd0 + lane], acc_p);
    if (lane >= 1) atomicAdd(&drel[dd0 + lane - 32], acc_n);
}

// ------------------------------------------------------------------------------------------------------------------
// forward: grid (ceil(Lq/128), H, B), 128 threads; thread t owns query row q0 + t (TMEM lane t)
// smem: Q 16K | K <=32K | V <=32K | P <=64K (k-blocks of 64 keys, 16K each) | brel | barriers
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) t5_attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                                                             const __grid_constant__ CUtensorMap tmv, TcArgs a, int lk_pad, int o_col,
                                                             int tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + smem_align_pad(smem_raw);      // 1024-byte aligned, still a __shared__ pointer (LDS / STS, 32-bit addressing)
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TILE * 128;
    uint8_t* sV = sK + lk_pad * 128;
    uint8_t* sP = sV + lk_pad * 128;
    const int nkb = (lk_pad + 63) / 64;
    float* brel = reinterpret_cast<float*>(sP + nkb * TILE * 128);           // [Lq + Lk]
    uint64_t* bars = reinterpret_cast<uint64_t*>(brel + ((a.Lq + a.Lk + 3) & ~3));
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

    const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z, bh = b * a.H + h;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int Lq = a.Lq, Lk = a.Lk;

    if (tid == 0) {
        mbar_init(&bars[0], 1);     // TMA
        mbar_init(&bars[1], 1);     // S ready
        mbar_init(&bars[2], 1);     // O ready
        fence_barrier_init();
        mbar_arrive_expect_tx(&bars[0], TILE * 128 + 2 * lk_pad * 128);
        tma_load_2d(sQ, &tmq, &bars[0], h * DK, b * Lq + q0);
        tma_load_2d(sK, &tmk, &bars[0], h * DK, b * Lk);
        tma_load_2d(sV, &tmv, &bars[0], h * DK, b * Lk);
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(tmem_ptr, tmem_cols);
        tmem_relinquish();
    }
    if (a.bias_table)
        for (int r = tid; r < Lq + Lk - 1; r += blockDim.x) {
            const int rel = r - (Lq - 1) - a.q_offset;
            brel[r] = a.bias_table[a.rel_bucket[rel + a.rel_zero] * a.H + h];
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (tid == 0) {
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(TILE, lk_pad, false, false);
        const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k)
            umma_bf16(tmem, umma_smem_desc_sw128(qa + k * 32, 16, 1024), umma_smem_desc_sw128(ka + k * 32, 16, 1024), idesc, k != 0);
        umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();

    const int i = q0 + tid;                                   // query row of this thread (may be >= Lq: computed, never stored)
    const int jmax = a.causal ? min(Lk, i + a.q_offset + 1) : Lk;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const uint64_t seed = a.seed + (a.seed_ptr ? *a.seed_ptr : 0ull);
    const DropKey dkey = make_drop_key(seed, a.dropout_p);
    const float* br = brel + (Lq - 1 - i);                     // br[j] = bias(j - i)
    const bool has_bias = a.bias_table != nullptr && i < Lq;

    // pass 1: row maximum
    float mx = -INFINITY;
    for (int c0 = 0; c0 < lk_pad; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(trow + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            const int j = c0 + t;
            if (j < jmax) mx = fmaxf(mx, __uint_as_float(r[t]) + (has_bias ? br[j] : 0.0f));
        }
    }
    // pass 2: p = exp(s - max), row sum, P (bf16, dropout applied) -> smem
    float sum = 0.0f;
    for (int c0 = 0; c0 < lk_pad; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(trow + c0, r);
        tmem_ld_wait();
        float p[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            const int j = c0 + t;
            float e = 0.0f;
            if (j < jmax) {
                e = __expf(__uint_as_float(r[t]) + (has_bias ? br[j] : 0.0f) - mx);
                sum += e;
                if (a.dropout_p > 0.0f) e *= drop_mult(a, dkey, bh, i, j);
            }
            p[t] = e;
        }
        uint8_t* blk = sP + (c0 >> 6) * (TILE * 128);
#pragma unroll
        for (int g = 0; g < 4; ++g) st_tile8(blk, tid, ((c0 & 63) >> 3) + g, p + 8 * g);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16(TILE, DK, false, true);
        const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
        for (int ks = 0; ks < lk_pad / 16; ++ks) {
            const uint64_t da = umma_smem_desc_sw128(pa + (ks >> 2) * (TILE * 128) + (ks & 3) * 32, 16, 1024);
            const uint64_t db = umma_smem_desc_sw128(va + ks * 2048, 8192, 1024);
            umma_bf16(tmem + o_col, da, db, idesc, ks != 0);
        }
        umma_commit(&bars[2]);
    }
    mbar_wait(&bars[2], 0);
    tc_fence_after();

    const float inv = 1.0f / sum;
#pragma unroll
    for (int c0 = 0; c0 < DK; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(trow + o_col + c0, r);
        tmem_ld_wait();
        if (i < Lq) {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + (static_cast<long long>(b) * Lq + i) * a.ldo + h * DK + c0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 q;
                q.x = pack_bf16(__uint_as_float(r[8 * g]) * inv, __uint_as_float(r[8 * g + 1]) * inv);
                q.y = pack_bf16(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv);
                q.z = pack_bf16(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv);
                q.w = pack_bf16(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv);
                reinterpret_cast<uint4*>(op)[g] = q;
            }
        }
    }
    if (i < Lq) a.lse[static_cast<long long>(bh) * Lq + i] = mx + __logf(sum);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward: grid (H, B), 128 threads.  Key blocks of 128 (kb) x query tiles of 128 (qt):
//   S = Q_qt K_kb^T, dP = dO_qt V_kb^T            -> TMEM (2 x 128 cols)
//   P = exp(S + bias - lse), dS = P (dP*m - D)    -> smem (bf16, two [128 x 64] swizzled blocks each)
//   dV_kb += P^T dO_qt, dK_kb += dS^T Q_qt        -> TMEM (2 x 64 cols), written out after the qt loop
//   dQ_qt += dS K_kb                              -> TMEM (2 x 64 cols), written out at the end
// smem: Q 2x16K | dO 2x16K | K 2x16K | V 2x16K | P 32K | dS 32K | brel/drel | barriers      (~200 KB, 1 CTA / SM)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) t5_attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                                                             const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
                                                             TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + smem_align_pad(smem_raw);      // 1024-byte aligned, still a __shared__ pointer (LDS / STS, 32-bit addressing)
    constexpr int T = TILE * 128;          // bytes of one [128 x 64] bf16 tile
    const int nq = (a.Lq + TILE - 1) / TILE, nk = (a.Lk + TILE - 1) / TILE;
    uint8_t* sQ = smem;
    uint8_t* sdO = sQ + nq * T;
    uint8_t* sK = sdO + nq * T;
    uint8_t* sV = sK + nk * T;
    uint8_t* sP = sV + nk * T;              // 2 blocks of 64 keys
    uint8_t* sdS = sP + 2 * T;
    float* brel = reinterpret_cast<float*>(sdS + 2 * T);                     // [Lq + Lk] bias per relative position
    float* drel = brel + ((a.Lq + a.Lk + 3) & ~3);                           // [Lq + Lk] dS summed per relative position
    float* bins = drel + ((a.Lq + a.Lk + 3) & ~3);                           // [num_buckets] per-bucket sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(bins + ((a.num_buckets + 3) & ~3));
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

    const int h = blockIdx.x, b = blockIdx.y, bh = b * a.H + h;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int Lq = a.Lq, Lk = a.Lk;
    constexpr int TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384;   // + 64 per query tile

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(&bars[0], 2 * (nq + nk) * T);
        for (int t = 0; t < nq; ++t) {
            tma_load_2d(sQ + t * T, &tmq, &bars[0], h * DK, b * Lq + t * TILE);
            tma_load_2d(sdO + t * T, &tmdo, &bars[0], h * DK, b * Lq + t * TILE);
        }
        for (int t = 0; t < nk; ++t) {
            tma_load_2d(sK + t * T, &tmk, &bars[0], h * DK, b * Lk + t * TILE);
            tma_load_2d(sV + t * T, &tmv, &bars[0], h * DK, b * Lk + t * TILE);
        }
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    const bool has_bias = a.bias_table != nullptr;
    if (has_bias)
        for (int r = tid; r < Lq + Lk - 1; r += blockDim.x) {
            const int rel = r - (Lq - 1) - a.q_offset;
            brel[r] = a.bias_table[a.rel_bucket[rel + a.rel_zero] * a.H + h];
            drel[r] = 0.0f;
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const uint64_t seed = a.seed + (a.seed_ptr ? *a.seed_ptr : 0ull);
    const DropKey dkey = make_drop_key(seed, a.dropout_p);

    // per query tile: D_i = dO_i . O_i and lse_i of the row this thread owns (global loads, overlapped with the TMA)
    float Dv[2] = {0.0f, 0.0f}, lsev[2] = {0.0f, 0.0f};
    for (int qt = 0; qt < nq; ++qt) {
        const int i = qt * TILE + tid;
        if (i < Lq) {
            const long long row = static_cast<long long>(b) * Lq + i;
            const uint4* dop = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.dout) + row * a.ldo + h * DK);
            const uint4* op = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.o) + row * a.ldo + h * DK);
            float acc = 0.0f;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const uint4 x = dop[g], y = op[g];
                const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&x);
                const __nv_bfloat162* yh = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float2 xf = __bfloat1622float2(xh[t]), yf = __bfloat1622float2(yh[t]);
                    acc = fmaf(xf.x, yf.x, fmaf(xf.y, yf.y, acc));
                }
            }
            Dv[qt] = acc;
            lsev[qt] = a.lse[static_cast<long long>(bh) * Lq + i];
        }
    }

    uint32_t phase = 0;
    if (tid == 0) {
        mbar_wait(&bars[0], 0);
        tc_fence_after();
    }
    __syncthreads();

    const uint32_t id_s = umma_idesc_bf16(TILE, TILE, false, false);     // S / dP : A K-major, B K-major, N = 128
    const uint32_t id_t = umma_idesc_bf16(TILE, DK, true, true);         // dV / dK: A MN-major (P^T), B MN-major, N = 64
    const uint32_t id_q = umma_idesc_bf16(TILE, DK, false, true);        // dQ     : A K-major (dS), B MN-major (K), N = 64

    for (int kb = 0; kb < nk; ++kb) {
        for (int qt = 0; qt < nq; ++qt) {
            if (tid == 0) {
                const uint32_t qa = smem_u32(sQ + qt * T), doa = smem_u32(sdO + qt * T);
                const uint32_t ka = smem_u32(sK + kb * T), va = smem_u32(sV + kb * T);
#pragma unroll
                for (int k = 0; k < DK / 16; ++k)
                    umma_bf16(tmem + TM_S, umma_smem_desc_sw128(qa + k * 32, 16, 1024), umma_smem_desc_sw128(ka + k * 32, 16, 1024), id_s, k != 0);
#pragma unroll
                for (int k = 0; k < DK / 16; ++k)
                    umma_bf16(tmem + TM_DP, umma_smem_desc_sw128(doa + k * 32, 16, 1024), umma_smem_desc_sw128(va + k * 32, 16, 1024), id_s, k != 0);
                umma_commit(&bars[1]);
            }
            mbar_wait(&bars[1], phase);
            phase ^= 1;
            tc_fence_after();

            const int i = qt * TILE + tid;
            const int jmax = a.causal ? min(Lk, i + a.q_offset + 1) : Lk;
            const float* br = brel + (Lq - 1 - i);
            const bool row_ok = i < Lq;
            for (int c0 = 0; c0 < TILE; c0 += 32) {
                uint32_t rs[32], rp[32];
                tmem_ld_32x32(trow + TM_S + c0, rs);
                tmem_ld_32x32(trow + TM_DP + c0, rp);
                tmem_ld_wait();
                float pv[32], dsv[32];
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const int j = kb * TILE + c0 + t;
                    float p = 0.0f, ds = 0.0f;
                    if (row_ok && j < jmax) {
                        const float pr = __expf(__uint_as_float(rs[t]) + (has_bias ? br[j] : 0.0f) - lsev[qt]);
                        const float m = a.dropout_p > 0.0f ? drop_mult(a, dkey, bh, i, j) : 1.0f;
                        p = pr * m;
                        ds = pr * (__uint_as_float(rp[t]) * m - Dv[qt]);
                    }
                    pv[t] = p;
                    dsv[t] = ds;
                }
                uint8_t* pb = sP + (c0 >> 6) * T;
                uint8_t* db = sdS + (c0 >> 6) * T;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    st_tile8(pb, tid, ((c0 & 63) >> 3) + g, pv + 8 * g);
                    st_tile8(db, tid, ((c0 & 63) >> 3) + g, dsv + 8 * g);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            if (tid == 0) {
                const uint32_t pa = smem_u32(sP), dsa = smem_u32(sdS);
                const uint32_t qa = smem_u32(sQ + qt * T), doa = smem_u32(sdO + qt * T), ka = smem_u32(sK + kb * T);
                // dV_kb += P^T dO_qt ; dK_kb += dS^T Q_qt : M = 128 keys (MN-major A: atoms of 64 keys are T bytes apart), K = 128 queries
#pragma unroll
                for (int ks = 0; ks < TILE / 16; ++ks) {
                    umma_bf16(tmem + TM_DV, umma_smem_desc_sw128(pa + ks * 2048, T, 1024), umma_smem_desc_sw128(doa + ks * 2048, 8192, 1024),
                              id_t, (qt | ks) != 0);
                    umma_bf16(tmem + TM_DK, umma_smem_desc_sw128(dsa + ks * 2048, T, 1024), umma_smem_desc_sw128(qa + ks * 2048, 8192, 1024),
                              id_t, (qt | ks) != 0);
                }
                // dQ_qt += dS K_kb : A K-major (k-blocks of 64 keys are T bytes apart), B = K_kb MN-major, K = 128 keys
#pragma unroll
                for (int ks = 0; ks < TILE / 16; ++ks)
                    umma_bf16(tmem + TM_DQ + qt * DK, umma_smem_desc_sw128(dsa + (ks >> 2) * T + (ks & 3) * 32, 16, 1024),
                              umma_smem_desc_sw128(ka + ks * 2048, 8192, 1024), id_q, (kb | ks) != 0);
                umma_commit(&bars[1]);
            }
            // bias gradient, overlapped with the MMAs: sum the dS tile along its diagonals (j - i = const).  Thread t owns
            // diagonals t and t + 128 of this tile, so the adds into drel[] are race free; at a fixed row the 32 lanes of a
            // warp read 32 consecutive bf16 of that row (no bank conflicts) -- replaces 128 shared atomics per thread.
            if (has_bias) {
                for (int half = 0; half < 2; ++half) {
                    const int dd = tid + half * TILE;                 // j_local - i_local + 127
                    if (dd > 2 * TILE - 2) break;
                    const int lo = max(0, TILE - 1 - dd), hi = min(TILE - 1, 2 * TILE - 2 - dd);
                    float acc4[4] = {0.0f, 0.0f, 0.0f, 0.0f};            // independent partial sums: the loads overlap
                    auto elem = [&](int il) -> float {
                        const int jl = il + dd - (TILE - 1);
                        return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(sdS + (jl >> 6) * T + sw128(il, jl & 63)));
                    };
                    int il = lo;
                    for (; il + 3 <= hi; il += 4) {
                        acc4[0] += elem(il);
                        acc4[1] += elem(il + 1);
                        acc4[2] += elem(il + 2);
                        acc4[3] += elem(il + 3);
                    }
                    for (; il <= hi; ++il) acc4[0] += elem(il);
                    const float acc = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
                    const int r = dd - (TILE - 1) + (kb - qt) * TILE + (Lq - 1);
                    if (r >= 0 && r < Lq + Lk - 1) drel[r] += acc;
                }
            }
            // the MMAs above read sP / sdS: wait for them before the next iteration overwrites the tiles
            mbar_wait(&bars[1], phase);
            phase ^= 1;
            tc_fence_after();
        }
        // dV_kb, dK_kb complete: lane t = key kb*128 + t
        const int j = kb * TILE + tid;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
#pragma unroll
            for (int c0 = 0; c0 < DK; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(trow + (which ? TM_DK : TM_DV) + c0, r);
                tmem_ld_wait();
                if (j < Lk) {
                    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(which ? a.dk : a.dv);
                    const long long ld = which ? a.ldk : a.ldv;
                    uint4* dst = reinterpret_cast<uint4*>(base + (static_cast<long long>(b) * Lk + j) * ld + h * DK + c0);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 q;
                        q.x = pack_bf16(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]));
                        q.y = pack_bf16(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
                        q.z = pack_bf16(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
                        q.w = pack_bf16(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
                        dst[g] = q;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();          // all lanes have drained dV / dK before the next key block re-initialises them
        tc_fence_after();
    }
    // dQ
    for (int qt = 0; qt < nq; ++qt) {
        const int i = qt * TILE + tid;
#pragma unroll
        for (int c0 = 0; c0 < DK; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(trow + TM_DQ + qt * DK + c0, r);
            tmem_ld_wait();
            if (i < Lq) {
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.dq) + (static_cast<long long>(b) * Lq + i) * a.ldq + h * DK + c0);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 q;
                    q.x = pack_bf16(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]));
                    q.y = pack_bf16(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
                    q.z = pack_bf16(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
                    q.w = pack_bf16(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
                    dst[g] = q;
                }
            }
        }
    }
    // bias gradient: per relative position -> per bucket partial of this (b, h)
    if (has_bias) {
        __syncthreads();
        float* part = a.dbias_partial + static_cast<long long>(bh) * a.num_buckets;
        for (int r = tid; r < a.num_buckets; r += blockDim.x) bins[r] = 0.0f;
        __syncthreads();
        for (int r = tid; r < Lq + Lk - 1; r += blockDim.x) {
            const int rel = r - (Lq - 1) - a.q_offset;
            atomicAdd(&bins[a.rel_bucket[rel + a.rel_zero]], drel[r]);
        }
        __syncthreads();
        for (int r = tid; r < a.num_buckets; r += blockDim.x) part[r] = bins[r];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Single-tile kernels (Lq <= 128 and Lk <= 128: every sequence of workloads 1-3): PERSISTENT, 512 threads.
//
// One (batch, head) problem is a short serial chain (TMA -> S [, dP] MMAs -> softmax -> P V / dV dK dQ MMAs -> write-out), so
// the kernels above -- one problem per CTA, 128 threads, one CTA per SM because of shared memory / TMEM -- are pure latency:
// ~22 us per backward problem for ~1 us of tensor work.  Here a CTA walks problems bh = blockIdx.x, + gridDim.x, ...:
//   * the operands of problem i+1 are fetched by TMA into a second buffer set while problem i computes;
//   * every SIMT phase is spread over 16 warps: 4 threads per query row (TMEM lane), 32 keys each; row statistics are
//     combined through shared memory;
//   * backward takes D_i = sum_j P~_ij dP_ij from the same P~ and dP that form dS (no dO . O re-read, exact cancellation);
//   * short self-attention problems are PACKED: with Lq = Lk = L in {32, 64} (decoder self-attention, the frozen text encoder)
//     a 128-row tile holds the G = 128 / L problems of consecutive batch entries of one head -- the TMA box already spans
//     them, the rows are contiguous in the projection output -- and S = Q K^T is used block-diagonally: row r belongs to
//     problem r / L, keys of other problems get P = dS = 0 exactly, so P V, dV, dK, dQ and the diagonal sums of the bias
//     gradient come out right without further masking.  A 32 x 32 problem used to pay for a whole 128 x 128 tile.
// ------------------------------------------------------------------------------------------------------------------
constexpr int ST_THREADS = 512;

__device__ __forceinline__ void st_tile16(uint8_t* tile, int row, int col16, const float* v) {        // 16 consecutive columns
    st_tile8(tile, row, col16 * 2, v);
    st_tile8(tile, row, col16 * 2 + 1, v + 8);
}

// smem: [Q 16K | K 16K | V 16K] x 2 | P 32K | brel | stats | barriers
__global__ void __launch_bounds__(ST_THREADS, 1) t5_attn_fwd_tc1_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                                                                        const __grid_constant__ CUtensorMap tmv, TcArgs a, int lk_pad) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + smem_align_pad(smem_raw);      // 1024-byte aligned, still a __shared__ pointer (LDS / STS, 32-bit addressing)
    constexpr int T = TILE * 128;
    uint8_t* sIn = smem;                          // 2 x (Q, K, V)
    uint8_t* sP = sIn + 6 * T;                    // 2 key blocks of 64
    float* brel = reinterpret_cast<float*>(sP + 2 * T);          // [256]
    float* sred = brel + 256;                                    // [4][128] row max partials
    float* ssum = sred + 4 * TILE;                               // [4][128] row sum partials (own array: no barrier between the two uses)
    uint64_t* bars = reinterpret_cast<uint64_t*>(ssum + 4 * TILE);   // full[2], mma
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = (warp & 3) * 32 + lane, cq = warp >> 2;       // TMEM lane = query row; key quarter
    const int Lq = a.Lq, Lk = a.Lk;
    const int G = a.pack;                                   // problems per tile (1: one problem, rows = queries)
    const int nprob = ((a.B + G - 1) / G) * a.H;            // work items: (batch group, head)
    constexpr int O_COL = 128;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(tmem_ptr, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint64_t seed = a.seed + (a.seed_ptr ? *a.seed_ptr : 0ull);
    const DropKey dkey = make_drop_key(seed, a.dropout_p);
    const uint32_t id_s = umma_idesc_bf16(TILE, TILE, false, false);
    const uint32_t id_o = umma_idesc_bf16(TILE, DK, false, true);
    const int nks = lk_pad / 16;

    auto issue_loads = [&](int bh, int buf) {
        const int bg = bh / a.H, h = bh - bg * a.H, b = bg * G;
        uint8_t* base = sIn + buf * 3 * T;
        mbar_arrive_expect_tx(&bars[buf], 3 * T);
        tma_load_2d(base, &tmq, &bars[buf], h * DK, b * Lq);
        tma_load_2d(base + T, &tmk, &bars[buf], h * DK, b * Lk);
        tma_load_2d(base + 2 * T, &tmv, &bars[buf], h * DK, b * Lk);
    };
    // problems are claimed dynamically (common.cuh: WorkClaim), one ahead of the one being computed so that its operands
    // can be prefetched; s_bh[buf] = problem whose operands sit in buffer set buf
    __shared__ int s_bh[2];
    WorkClaim wc;
    wc.init(a.sched);
    int n1 = nprob, n2 = nprob;                            // thread 0: the next problem and the one after (claimed ahead: the
    if (tid == 0) {                                        // atomic's round trip must not sit in front of an MMA issue)
        const int first = wc.first();
        s_bh[0] = first;
        if (first < nprob) {
            issue_loads(first, 0);
            n1 = wc.claim();
        }
    }
    __syncthreads();

    uint32_t mma_phase = 0;
    int it = 0;
    for (int bh = s_bh[0]; bh < nprob; ++it) {
        const int buf = it & 1;
        const int h = bh % a.H, b = (bh / a.H) * G;           // b: first batch entry of the tile
        uint8_t* sQ = sIn + buf * 3 * T;
        uint8_t* sK = sQ + T;
        uint8_t* sV = sK + T;
        if (tid == 0) {
            s_bh[buf ^ 1] = n1;
            if (n1 < nprob) {
                issue_loads(n1, buf ^ 1);
                n2 = wc.claim();                           // consumed at the end of this iteration
            } else {
                wc.finish_begin();                         // this is the CTA's last iteration (bh becomes n1 >= nprob): its claims
                n2 = nprob;                                //   are over; the atomic's round trip overlaps the last problem
            }
        }
        const bool has_bias = a.bias_table != nullptr;
        if (has_bias && tid < Lq + Lk - 1) {
            const int rel = tid - (Lq - 1) - a.q_offset;
            brel[tid] = a.bias_table[a.rel_bucket[rel + a.rel_zero] * a.H + h];
        }
        if (tid == 0) {
            mbar_wait(&bars[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
            for (int k = 0; k < DK / 16; ++k)
                umma_bf16(tmem, umma_smem_desc_sw128(qa + k * 32, 16, 1024), umma_smem_desc_sw128(ka + k * 32, 16, 1024), id_s, k != 0);
            umma_commit(&bars[2]);
        }
        __syncthreads();                                   // brel visible
        mbar_wait(&bars[2], mma_phase);
        mma_phase ^= 1;
        tc_fence_after();

        // row r of the tile = query i of sub-problem `sub` (packed) or query r of the only problem
        const int sub = G > 1 ? r / Lq : 0;
        const int i = r - sub * Lq;
        const bool row_ok = G > 1 ? b + sub < a.B : i < Lq;
        const long long bh_row = static_cast<long long>(b + sub) * a.H + h;
        const int jmax = a.causal ? min(Lk, i + a.q_offset + 1) : Lk;
        const int j0 = cq * 32;                            // first of this thread's 32 keys (tile column)
        const bool own = G == 1 || j0 / Lk == sub;         // the 32 keys lie inside one problem (Lk is 32 or 64 when packed)
        const int jb = j0 - sub * Lk;                      // ... at key index jb of it
        const float* br = brel + (Lq - 1 - i) + jb;
        const bool bias_on = has_bias && row_ok && own;
        // warp-uniform: the warp's 32 keys exist, and at least one of its 32 rows is a query (cross-attention with 32 targets: one
        // row-warp in four; the others used to run both softmax passes on padding rows)
        const bool active = j0 < lk_pad && (warp & 3) * 32 < (G > 1 ? TILE : Lq);
        float sv[32];
        float mx = -INFINITY;
        if (active) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t rr[16];
                tmem_ld_32x16(trow + j0 + hf * 16, rr);
                tmem_ld_wait();
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const int j = jb + hf * 16 + t;
                    float sc = -INFINITY;
                    if (own && j < jmax) {
                        sc = __uint_as_float(rr[t]) + (bias_on ? br[hf * 16 + t] : 0.0f);
                        mx = fmaxf(mx, sc);
                    }
                    sv[hf * 16 + t] = sc;
                }
            }
        }
        sred[cq * TILE + r] = mx;
        __syncthreads();
        mx = fmaxf(fmaxf(sred[r], sred[TILE + r]), fmaxf(sred[2 * TILE + r], sred[3 * TILE + r]));
        float sum = 0.0f;
        if (active) {
            const uint64_t base = (static_cast<uint64_t>(bh_row) * Lq + i) * Lk + jb;
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                const float e = sv[t] > -INFINITY ? __expf(sv[t] - mx) : 0.0f;
                sum += e;
                sv[t] = e;
            }
            if (dkey.on && row_ok && own) dropout_apply_run<32>(dkey, base, sv);     // elements past jmax are zero already
            uint8_t* blk = sP + (j0 >> 6) * T;
            st_tile16(blk, r, (j0 & 63) >> 4, sv);
            st_tile16(blk, r, ((j0 & 63) >> 4) + 1, sv + 16);
        }
        ssum[cq * TILE + r] = sum;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
            for (int ks = 0; ks < nks; ++ks)
                umma_bf16(tmem + O_COL, umma_smem_desc_sw128(pa + (ks >> 2) * T + (ks & 3) * 32, 16, 1024),
                          umma_smem_desc_sw128(va + ks * 2048, 8192, 1024), id_o, ks != 0);
            umma_commit(&bars[2]);
        }
        sum = (ssum[r] + ssum[TILE + r]) + (ssum[2 * TILE + r] + ssum[3 * TILE + r]);
        mbar_wait(&bars[2], mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        if (cq < 2 && (warp & 3) * 32 < (G > 1 ? TILE : Lq)) {      // warps 0-3: output columns 0-31, warps 4-7: 32-63 (row-warps with queries)
            uint32_t ro[32];
            tmem_ld_32x32(trow + O_COL + cq * 32, ro);
            tmem_ld_wait();
            if (row_ok) {
                const float inv = 1.0f / sum;
                __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + (static_cast<long long>(b) * Lq + r) * a.ldo + h * DK + cq * 32;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    uint4 q;
                    q.x = pack_bf16(__uint_as_float(ro[8 * gq]) * inv, __uint_as_float(ro[8 * gq + 1]) * inv);
                    q.y = pack_bf16(__uint_as_float(ro[8 * gq + 2]) * inv, __uint_as_float(ro[8 * gq + 3]) * inv);
                    q.z = pack_bf16(__uint_as_float(ro[8 * gq + 4]) * inv, __uint_as_float(ro[8 * gq + 5]) * inv);
                    q.w = pack_bf16(__uint_as_float(ro[8 * gq + 6]) * inv, __uint_as_float(ro[8 * gq + 7]) * inv);
                    reinterpret_cast<uint4*>(op)[gq] = q;
                }
            }
        } else if (cq == 2 && row_ok) {
            a.lse[bh_row * Lq + i] = mx + __logf(sum);
        }
        tc_fence_before();
        __syncthreads();                                   // TMEM, sP, brel, sred are reused by the next problem
        tc_fence_after();
        bh = s_bh[buf ^ 1];
        n1 = n2;
    }
    if (tid == 0) wc.finish_end();                         // (the grid never exceeds the problem count: every CTA ran >= 1 iteration)
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

// smem: [Q | dO | K | V] x 2 (16K each) | P 32K | dS 32K | brel | drel | bins | stats | barriers   (~197 KB)
// TMEM: S 0..127 | dP 128..255 | dV 256..319 | dK 320..383 | dQ 384..447
__global__ void __launch_bounds__(ST_THREADS, 1) t5_attn_bwd_tc1_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                                                                        const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
                                                                        TcArgs a, int lq_pad, int lk_pad) {
    extern __shared__ uint8_t smem_raw[];
    // (generic pointer on purpose: with a true __shared__ pointer this kernel measured 25 % SLOWER, twice -- 89.8 -> 114 us at 96 x 96 --
    // while the three other kernels of this file gained 3-10 % from LDS / STS addressing)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int T = TILE * 128;
    uint8_t* sIn = smem;                          // 2 x (Q, dO, K, V)
    uint8_t* sP = sIn + 8 * T;                    // 2 key blocks of 64
    uint8_t* sdS = sP + 2 * T;
    float* brel = reinterpret_cast<float*>(sdS + 2 * T);         // [256] bias per relative position
    float* drel = brel + 256;                                    // [256] dS summed per relative position
    float* bins = drel + 256;                                    // [64]
    float* sred = bins + 64;                                     // [4][128] partial D_i
    uint64_t* bars = reinterpret_cast<uint64_t*>(sred + 4 * TILE);   // full[2], mma
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
    constexpr int TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = (warp & 3) * 32 + lane, cq = warp >> 2;
    const int Lq = a.Lq, Lk = a.Lk;
    const int G = a.pack;                                   // problems per tile (see the forward kernel)
    const int nprob = ((a.B + G - 1) / G) * a.H;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    // rows of P / dS at and beyond lq_pad are never written and never read (the contraction over queries stops at lq_pad);
    // rows in [Lq, lq_pad) are written as zeros below
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint64_t seed = a.seed + (a.seed_ptr ? *a.seed_ptr : 0ull);
    const DropKey dkey = make_drop_key(seed, a.dropout_p);
    const uint32_t id_s = umma_idesc_bf16(TILE, TILE, false, false);     // S / dP : A K-major, B K-major, N = 128
    const uint32_t id_t = umma_idesc_bf16(TILE, DK, true, true);         // dV / dK: A MN-major (P^T), B MN-major, N = 64
    const uint32_t id_q = umma_idesc_bf16(TILE, DK, false, true);        // dQ     : A K-major (dS), B MN-major (K), N = 64
    const int nqs = lq_pad / 16, nks = lk_pad / 16;
    const bool has_bias = a.bias_table != nullptr;

    auto issue_loads = [&](int bh, int buf) {
        const int bg = bh / a.H, h = bh - bg * a.H, b = bg * G;
        uint8_t* base = sIn + buf * 4 * T;
        mbar_arrive_expect_tx(&bars[buf], 4 * T);
        tma_load_2d(base, &tmq, &bars[buf], h * DK, b * Lq);
        tma_load_2d(base + T, &tmdo, &bars[buf], h * DK, b * Lq);
        tma_load_2d(base + 2 * T, &tmk, &bars[buf], h * DK, b * Lk);
        tma_load_2d(base + 3 * T, &tmv, &bars[buf], h * DK, b * Lk);
    };
    __shared__ int s_bh[2];                                // see the forward kernel
    WorkClaim wc;
    wc.init(a.sched);
    int n1 = nprob, n2 = nprob;                            // thread 0: the next problem and the one after (claimed ahead: the
    if (tid == 0) {                                        // atomic's round trip must not sit in front of an MMA issue)
        const int first = wc.first();
        s_bh[0] = first;
        if (first < nprob) {
            issue_loads(first, 0);
            n1 = wc.claim();
        }
    }
    __syncthreads();

    uint32_t mma_phase = 0;
    int it = 0;
    for (int bh = s_bh[0]; bh < nprob; ++it) {
        const int buf = it & 1;
        const int h = bh % a.H, b = (bh / a.H) * G;           // b: first batch entry of the tile
        uint8_t* sQ = sIn + buf * 4 * T;
        uint8_t* sdO = sQ + T;
        uint8_t* sK = sdO + T;
        uint8_t* sV = sK + T;
        if (tid == 0) {
            s_bh[buf ^ 1] = n1;
            if (n1 < nprob) {
                issue_loads(n1, buf ^ 1);
                n2 = wc.claim();                           // consumed at the end of this iteration
            } else {
                wc.finish_begin();                         // this is the CTA's last iteration (bh becomes n1 >= nprob): its claims
                n2 = nprob;                                //   are over; the atomic's round trip overlaps the last problem
            }
        }
        if (has_bias && tid < Lq + Lk - 1) {
            const int rel = tid - (Lq - 1) - a.q_offset;
            brel[tid] = a.bias_table[a.rel_bucket[rel + a.rel_zero] * a.H + h];
        }
        if (tid < 256) drel[tid] = 0.0f;                      // dS summed per diagonal (tile column - tile row + 127)
        if (tid < 64) bins[tid] = 0.0f;
        const int sub = G > 1 ? r / Lq : 0;                   // row r = query i of sub-problem `sub`
        const int i = r - sub * Lq;
        const bool row_ok = G > 1 ? b + sub < a.B : i < Lq;
        const long long bh_row = static_cast<long long>(b + sub) * a.H + h;
        const float lse = row_ok ? a.lse[bh_row * Lq + i] : 0.0f;
        if (tid == 0) {
            mbar_wait(&bars[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t qa = smem_u32(sQ), doa = smem_u32(sdO), ka = smem_u32(sK), va = smem_u32(sV);
#pragma unroll
            for (int k = 0; k < DK / 16; ++k)
                umma_bf16(tmem + TM_S, umma_smem_desc_sw128(qa + k * 32, 16, 1024), umma_smem_desc_sw128(ka + k * 32, 16, 1024), id_s, k != 0);
#pragma unroll
            for (int k = 0; k < DK / 16; ++k)
                umma_bf16(tmem + TM_DP, umma_smem_desc_sw128(doa + k * 32, 16, 1024), umma_smem_desc_sw128(va + k * 32, 16, 1024), id_s, k != 0);
            umma_commit(&bars[2]);
        }
        __syncthreads();                                   // brel / drel / bins visible
        mbar_wait(&bars[2], mma_phase);
        mma_phase ^= 1;
        tc_fence_after();

        const int j0 = cq * 32;                            // first of this thread's 32 keys (tile column)
        const bool own = G == 1 || j0 / Lk == sub;         // they lie inside one problem (Lk is 32 or 64 when packed) ...
        const int jb = j0 - sub * Lk;                      // ... at key index jb of it
        const int jmax = own ? (a.causal ? min(Lk, i + a.q_offset + 1) : Lk) : -(1 << 30);      // foreign keys: P = dS = 0
        const float* br = brel + (Lq - 1 - i) + jb;
        const bool active = j0 < lk_pad && (warp & 3) * 32 < lq_pad;      // warp-uniform: keys exist and the warp owns rows of the P / dS tiles
        // pass 1: P~ = P * dropout multiplier (what the forward multiplied into V) and the partial D_i = sum_j P~_ij dP_ij
        uint32_t keep = 0xFFFFFFFFu;                       // dropout keep bits of the 32 keys
        uint8_t* pb = sP + (j0 >> 6) * T;
        uint8_t* db = sdS + (j0 >> 6) * T;
        const int c16 = (j0 & 63) >> 4;
        float Dp = 0.0f;
        if (active) {
            if (dkey.on) {
                const uint64_t base = (static_cast<uint64_t>(bh_row) * Lq + i) * Lk + (own ? jb : 0);
                keep = 0;
                if ((base & 1) == 0) {
#pragma unroll
                    for (int t = 0; t < 32; t += 2) {
                        const uint32_t hsh = drop_hash_pair(dkey, (base + t) >> 1);
                        keep |= ((hsh & 0xFFFFu) < dkey.thr16 ? 1u : 0u) << t;
                        keep |= ((hsh >> 16) < dkey.thr16 ? 1u : 0u) << (t + 1);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 32; ++t) keep |= (dropout_mult(dkey, base + t) != 0.0f ? 1u : 0u) << t;
                }
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t rs[16], rp[16];
                tmem_ld_32x16(trow + TM_S + j0 + hf * 16, rs);
                tmem_ld_32x16(trow + TM_DP + j0 + hf * 16, rp);
                tmem_ld_wait();
                float pm[16];                              // P~
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const int j = jb + hf * 16 + t;
                    float p = 0.0f;
                    if (row_ok && j < jmax) {
                        p = __expf(__uint_as_float(rs[t]) + (has_bias ? br[hf * 16 + t] : 0.0f) - lse);
                        p = (keep >> (hf * 16 + t)) & 1u ? p * dkey.inv_keep : 0.0f;
                        Dp = fmaf(p, __uint_as_float(rp[t]), Dp);
                    }
                    pm[t] = p;
                }
                if (r < lq_pad) st_tile16(pb, r, c16 + hf, pm);
            }
        }
        sred[cq * TILE + r] = Dp;
        __syncthreads();
        const float Di = (sred[r] + sred[TILE + r]) + (sred[2 * TILE + r] + sred[3 * TILE + r]);
        // pass 2: dS_ij = P_ij (m_ij dP_ij - D_i) with P = P~ / m on kept elements; dropped elements keep P_ij D_i, which needs the
        // un-dropped probability -> recompute it from S for those (rare: p = 0.1)
        if (active) {                                       // warp-uniform (tcgen05.ld is warp-collective); stores are per lane
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float dsv[16];
                uint32_t rs[16], rp[16];
                tmem_ld_32x16(trow + TM_S + j0 + hf * 16, rs);
                tmem_ld_32x16(trow + TM_DP + j0 + hf * 16, rp);
                tmem_ld_wait();
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const int j = jb + hf * 16 + t;
                    float ds = 0.0f;
                    if (row_ok && j < jmax) {
                        const bool kept = (keep >> (hf * 16 + t)) & 1u;
                        const float pr = __expf(__uint_as_float(rs[t]) + (has_bias ? br[hf * 16 + t] : 0.0f) - lse);
                        ds = pr * ((kept ? __uint_as_float(rp[t]) * dkey.inv_keep : 0.0f) - Di);
                    }
                    dsv[t] = ds;
                }
                if (r < lq_pad) st_tile16(db, r, c16 + hf, dsv);
                if (has_bias) diag_accumulate16(drel, dsv, j0 + hf * 16 - (warp & 3) * 32 + (TILE - 1), lane);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            const uint32_t pa = smem_u32(sP), dsa = smem_u32(sdS);
            const uint32_t qa = smem_u32(sQ), doa = smem_u32(sdO), ka = smem_u32(sK);
            // dV = P~^T dO ; dK = dS^T Q : M = 128 keys (MN-major A: atoms of 64 keys are T bytes apart), contraction over lq_pad queries
            for (int ks = 0; ks < nqs; ++ks) {
                umma_bf16(tmem + TM_DV, umma_smem_desc_sw128(pa + ks * 2048, T, 1024), umma_smem_desc_sw128(doa + ks * 2048, 8192, 1024), id_t, ks != 0);
                umma_bf16(tmem + TM_DK, umma_smem_desc_sw128(dsa + ks * 2048, T, 1024), umma_smem_desc_sw128(qa + ks * 2048, 8192, 1024), id_t, ks != 0);
            }
            // dQ = dS K : A K-major (k-blocks of 64 keys are T bytes apart), B = K MN-major, contraction over lk_pad keys
            for (int ks = 0; ks < nks; ++ks)
                umma_bf16(tmem + TM_DQ, umma_smem_desc_sw128(dsa + (ks >> 2) * T + (ks & 3) * 32, 16, 1024),
                          umma_smem_desc_sw128(ka + ks * 2048, 8192, 1024), id_q, ks != 0);
            umma_commit(&bars[2]);
        }
        // bias gradient, overlapped with the MMAs: diagonal sums (formed in pass 2, see diag_accumulate16) -> buckets
        if (has_bias && tid < 2 * TILE - 1) {
            const int rr = tid - (TILE - 1) + (Lq - 1);       // relative position j - i + (Lq - 1) of diagonal `tid`
            const float acc = drel[tid];
            if (rr >= 0 && rr < Lq + Lk - 1 && acc != 0.0f) atomicAdd(&bins[a.rel_bucket[rr - (Lq - 1) - a.q_offset + a.rel_zero]], acc);
        }
        mbar_wait(&bars[2], mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        // write-out: warps 0-3 dq, 4-7 dk, 8-11 dv (row = TMEM lane)
        if (cq < 3) {
            const int rows = cq == 0 ? Lq : Lk;
            const int rows_ok = G > 1 ? min(TILE, (a.B - b) * rows) : rows;      // valid rows of the tile
            const uint32_t col = cq == 0 ? TM_DQ : (cq == 1 ? TM_DK : TM_DV);
            __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(cq == 0 ? a.dq : (cq == 1 ? a.dk : a.dv));
            const long long ld = cq == 0 ? a.ldq : (cq == 1 ? a.ldk : a.ldv);
#pragma unroll
            for (int c0 = 0; c0 < DK; c0 += 32) {
                uint32_t ro[32];
                tmem_ld_32x32(trow + col + c0, ro);
                tmem_ld_wait();
                if (r < rows_ok) {
                    uint4* dst = reinterpret_cast<uint4*>(base + (static_cast<long long>(b) * rows + r) * ld + h * DK + c0);
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                        uint4 q;
                        q.x = pack_bf16(__uint_as_float(ro[8 * gq]), __uint_as_float(ro[8 * gq + 1]));
                        q.y = pack_bf16(__uint_as_float(ro[8 * gq + 2]), __uint_as_float(ro[8 * gq + 3]));
                        q.z = pack_bf16(__uint_as_float(ro[8 * gq + 4]), __uint_as_float(ro[8 * gq + 5]));
                        q.w = pack_bf16(__uint_as_float(ro[8 * gq + 6]), __uint_as_float(ro[8 * gq + 7]));
                        dst[gq] = q;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();                                   // TMEM, sP / sdS, brel / bins are reused by the next problem
        tc_fence_after();
        // partial bias gradient of the tile: into the slot of its first (batch, head); the other packed problems' slots get zeros
        if (has_bias && tid < a.num_buckets) {
            for (int g = 0; g < G && b + g < a.B; ++g)
                a.dbias_partial[(static_cast<long long>(b + g) * a.H + h) * a.num_buckets + tid] = g == 0 ? bins[tid] : 0.0f;
        }
        bh = s_bh[buf ^ 1];
        n1 = n2;
    }
    if (tid == 0) wc.finish_end();                         // (the grid never exceeds the problem count: every CTA ran >= 1 iteration)
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Single-tile BACKWARD kernel, TWO CTAs PER SM (256 threads each).
//
// The 512-thread kernels above keep one problem per SM in flight, and one problem is a serial chain of latencies (TMA -> MMA ->
// commit -> tcgen05.ld -> softmax -> st.shared -> fence -> MMA -> ...): profiles/r02_attn_ncu.txt shows >= 55 % of their warp
// samples at CTA barriers / mbarrier waits.  This variant halves every per-CTA resource so that two CTAs are resident per SM
// and one CTA's chain fills the other's bubbles:
//   * 256 threads: TMEM lane (row) = (warp & 3) * 32 + lane, key half = warp >> 2, up to 64 keys per thread walked in 16-key
//     chunks (chunks interleaved between the halves: 96 keys split 3 + 3); both passes read S and dP from TMEM;
//   * operand tiles are as tall as the problem (rows padded to 16, not to 128) and SINGLE buffered; the loads of the CTA's next
//     problem are issued the moment the gradient MMAs have retired, so they travel under the write-out and the next preamble;
//     UMMA A operands always span 128 rows: rows past the tile are whatever follows in shared memory (the layout keeps those
//     reads inside the allocation); they only reach TMEM lanes that are never read (an MMA row depends on its own A row only);
//   * 256 TMEM columns per CTA: dV | dK | dQ overwrite S | dP once every thread has consumed them.
// Measured (B = 64, H = 16, us per launch, this kernel | 512-thread kernel): 96 x 96 self-attention 72.2 | 82.1, 32 x 96
// cross-attention 52.8 | 66.2 (profiles/r02_t5_attn_2cta.txt).  The same scheme for the FORWARD kernel was built and measured at
// parity or worse (45.4 | 45.3, 41.1 | 42.1, 128 x 128: 57.4 | 53.8) and removed: with the same number of resident warps each
// thread's serial share of a problem doubles, which cancels the overlap.  Used when the footprint allows two CTAs per SM
// (<= 113 KiB: not the packed 128-row tiles); KLAB_T5_ATTN_2CTA=0 switches it off.
// ------------------------------------------------------------------------------------------------------------------
constexpr int S2_THREADS = 256;
constexpr int S2_MAX_SMEM = 113 * 1024;

struct S2Layout {
    int tq, tk, nkb;          // bytes of a query-side tile (lq_pad rows), of a key-side tile (lk_pad rows), 64-key blocks of P / dS
    int q, dO, dS, p, k, v;   // byte offsets of the tiles
    int fl;                   // float scratch
    int total;                // bytes to allocate (without the 1 KiB alignment slack)
};

__host__ __device__ inline S2Layout s2_layout(int lq_pad, int lk_pad) {
    S2Layout L;
    L.tq = lq_pad * 128;
    L.tk = lk_pad * 128;
    L.nkb = (lk_pad + 63) / 64;
    // K-major A operands (128 rows are read whatever the tile height) first, so that the over-read lands in the tiles behind them
    L.q = 0;                                                 // A of S = Q K^T
    L.dO = L.q + L.tq;                                       // A of dP = dO V^T
    L.dS = L.dO + L.tq;                                      // A of dQ = dS K: block b is read up to b * tq + 16 KiB
    L.p = L.dS + L.nkb * L.tq;
    L.k = L.p + L.nkb * L.tq;
    L.v = L.k + L.tk;
    L.fl = L.v + L.tk;
    int o = L.fl + static_cast<int>(sizeof(float)) * (256 + 64 + 4 * TILE) + 128;    // brel | bins | [2][128] partial D (+ spare) | barriers, tmem pointer, next item
    // P^T / dS^T are MN-major A operands with M = 128 keys = two 64-key blocks even when only one is stored: the second then reads
    // the tq bytes behind the first
    const int need1 = L.dS + (L.nkb - 1) * L.tq + TILE * 128, need2 = L.p + 2 * L.tq, need3 = L.dO + TILE * 128;
    if (o < need1) o = need1;
    if (o < need2) o = need2;
    if (o < need3) o = need3;
    L.total = o;
    return L;
}

// TMEM: S 0.. | dP 128..  ->  dV 0..63 | dK 64..127 | dQ 128..191 (overwrite S / dP after the softmax backward)
__global__ void __launch_bounds__(S2_THREADS, 2) t5_attn_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                                                                        const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
                                                                        TcArgs a, int lq_pad, int lk_pad) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + smem_align_pad(smem_raw);
    const S2Layout L = s2_layout(lq_pad, lk_pad);
    uint8_t* sQ = smem + L.q;
    uint8_t* sdO = smem + L.dO;
    uint8_t* sdS = smem + L.dS;
    uint8_t* sP = smem + L.p;
    uint8_t* sK = smem + L.k;
    uint8_t* sV = smem + L.v;
    float* brel = reinterpret_cast<float*>(smem + L.fl);         // [256] bias per relative position
    float* bins = brel + 256;                                    // [64]
    float* sred = bins + 64;                                     // [2][128] partial D_i
    float* drel = sred + 2 * TILE;                               // [256] dS summed per diagonal (tile column - tile row + 127)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sred + 4 * TILE);   // operands full, mma
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);
    int* s_next = reinterpret_cast<int*>(tmem_ptr + 1);
    constexpr int TM_S = 0, TM_DP = 128, TM_DV = 0, TM_DK = 64, TM_DQ = 128;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = (warp & 3) * 32 + lane, ch = warp >> 2;
    const int Lq = a.Lq, Lk = a.Lk;
    const int G = a.pack;
    const int nprob = ((a.B + G - 1) / G) * a.H;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(tmem_ptr, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint64_t seed = a.seed + (a.seed_ptr ? *a.seed_ptr : 0ull);
    const DropKey dkey = make_drop_key(seed, a.dropout_p);
    const uint32_t id_s = umma_idesc_bf16(TILE, lk_pad, false, false);   // S / dP : A K-major, B K-major, N = lk_pad
    const uint32_t id_t = umma_idesc_bf16(TILE, DK, true, true);         // dV / dK: A MN-major (P^T), B MN-major, N = 64
    const uint32_t id_q = umma_idesc_bf16(TILE, DK, false, true);        // dQ     : A K-major (dS), B MN-major (K), N = 64
    const int nqs = lq_pad / 16, nks = lk_pad / 16;
    const bool has_bias = a.bias_table != nullptr;

    auto issue_loads = [&](int bh) {
        const int bg = bh / a.H, h = bh - bg * a.H, b = bg * G;
        mbar_arrive_expect_tx(&bars[0], 2 * L.tq + 2 * L.tk);
        tma_load_2d(sQ, &tmq, &bars[0], h * DK, b * Lq);
        tma_load_2d(sdO, &tmdo, &bars[0], h * DK, b * Lq);
        tma_load_2d(sK, &tmk, &bars[0], h * DK, b * Lk);
        tma_load_2d(sV, &tmv, &bars[0], h * DK, b * Lk);
    };
    WorkClaim wc;
    wc.init(a.sched);
    int n1 = nprob, n2 = nprob;
    if (tid == 0) {
        const int first = wc.first();
        *s_next = first;
        if (first < nprob) {
            issue_loads(first);
            n1 = wc.claim();
        }
    }
    __syncthreads();

    uint32_t mma_phase = 0;
    int it = 0;
    for (int bh = *s_next; bh < nprob; ++it) {
        const int h = bh % a.H, b = (bh / a.H) * G;
        if (has_bias && tid < Lq + Lk - 1) {
            const int rel = tid - (Lq - 1) - a.q_offset;
            brel[tid] = a.bias_table[a.rel_bucket[rel + a.rel_zero] * a.H + h];
        }
        if (tid < 64) bins[tid] = 0.0f;
        drel[tid] = 0.0f;                                     // (256 threads)
        const int sub = G > 1 ? r / Lq : 0;                   // row r = query i of sub-problem `sub`
        const int i = r - sub * Lq;
        const bool row_ok = G > 1 ? b + sub < a.B : i < Lq;
        const long long bh_row = static_cast<long long>(b + sub) * a.H + h;
        const float lse = row_ok ? a.lse[bh_row * Lq + i] : 0.0f;
        if (tid == 0) {
            if (n1 < nprob) {
                n2 = wc.claim();
            } else {
                wc.finish_begin();
                n2 = nprob;
            }
            mbar_wait(&bars[0], it & 1);
            tc_fence_after();
            const uint32_t qa = smem_u32(sQ), doa = smem_u32(sdO), ka = smem_u32(sK), va = smem_u32(sV);
#pragma unroll
            for (int k = 0; k < DK / 16; ++k)
                umma_bf16(tmem + TM_S, umma_smem_desc_sw128(qa + k * 32, 16, 1024), umma_smem_desc_sw128(ka + k * 32, 16, 1024), id_s, k != 0);
#pragma unroll
            for (int k = 0; k < DK / 16; ++k)
                umma_bf16(tmem + TM_DP, umma_smem_desc_sw128(doa + k * 32, 16, 1024), umma_smem_desc_sw128(va + k * 32, 16, 1024), id_s, k != 0);
            umma_commit(&bars[1]);
        }
        __syncthreads();                                   // brel / bins visible; every thread has read s_next (= bh)
        mbar_wait(&bars[1], mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        if (tid == 0) *s_next = n1;                        // read after the last barrier of this iteration

        const int jmax_p = a.causal ? min(Lk, i + a.q_offset + 1) : Lk;
        const float* br0 = brel + (Lq - 1 - i);
        const bool lanes_live = (warp & 3) * 32 < lq_pad;  // warp-uniform: this warp owns at least one row of the P / dS tiles
        // pass 1: P~ = P * dropout multiplier (what the forward multiplied into V) and the partial D_i = sum_j P~_ij dP_ij
        unsigned long long keep = ~0ull;                   // dropout keep bits of this thread's 64 keys
        float Dp = 0.0f;
#pragma unroll 1
        for (int c = 0; c < (lanes_live ? 4 : 0); ++c) {
            const int j0 = (2 * c + ch) * 16;                // tile column of the chunk (warp-uniform); halves interleaved: balanced for 96 keys
            if (j0 >= lk_pad) break;
            const bool own = G == 1 || j0 / Lk == sub;      // the 16 keys lie inside one problem (Lk is 32 or 64 when packed) ...
            const int jb = j0 - sub * Lk;                   // ... at key index jb of it
            const int jmax = own ? jmax_p : -(1 << 30);     // foreign keys: P = dS = 0
            uint32_t kbits = 0xFFFFu;
            if (dkey.on) {
                const uint64_t base = (static_cast<uint64_t>(bh_row) * Lq + i) * Lk + (own ? jb : 0);
                kbits = 0;
                if ((base & 1) == 0) {
#pragma unroll
                    for (int t = 0; t < 16; t += 2) {
                        const uint32_t hsh = drop_hash_pair(dkey, (base + t) >> 1);
                        kbits |= ((hsh & 0xFFFFu) < dkey.thr16 ? 1u : 0u) << t;
                        kbits |= ((hsh >> 16) < dkey.thr16 ? 1u : 0u) << (t + 1);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 16; ++t) kbits |= (dropout_mult(dkey, base + t) != 0.0f ? 1u : 0u) << t;
                }
                keep = (keep & ~(0xFFFFull << (c * 16))) | (static_cast<unsigned long long>(kbits) << (c * 16));
            }
            uint32_t rs[16], rp[16];
            tmem_ld_32x16(trow + TM_S + j0, rs);
            tmem_ld_32x16(trow + TM_DP + j0, rp);
            tmem_ld_wait();
            float pm[16];                                  // P~
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int j = jb + t;
                float p = 0.0f;
                if (row_ok && j < jmax) {
                    p = __expf(__uint_as_float(rs[t]) + (has_bias ? br0[j] : 0.0f) - lse);
                    p = (kbits >> t) & 1u ? p * dkey.inv_keep : 0.0f;
                    Dp = fmaf(p, __uint_as_float(rp[t]), Dp);
                }
                pm[t] = p;
            }
            if (r < lq_pad) st_tile16(sP + (j0 >> 6) * L.tq, r, (j0 & 63) >> 4, pm);
        }
        sred[ch * TILE + r] = Dp;
        __syncthreads();
        const float Di = sred[r] + sred[TILE + r];
        // pass 2: dS_ij = P_ij (m_ij dP_ij - D_i); dropped elements keep -P_ij D_i with the UN-dropped probability, recomputed from S
        if (lanes_live) {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int j0 = (2 * c + ch) * 16;
                if (j0 >= lk_pad) break;
                const bool own = G == 1 || j0 / Lk == sub;
                const int jb = j0 - sub * Lk;
                const int jmax = own ? jmax_p : -(1 << 30);
                const uint32_t kbits = static_cast<uint32_t>(keep >> (c * 16)) & 0xFFFFu;
                uint32_t rs[16], rp[16];
                tmem_ld_32x16(trow + TM_S + j0, rs);
                tmem_ld_32x16(trow + TM_DP + j0, rp);
                tmem_ld_wait();
                float dsv[16];
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const int j = jb + t;
                    float ds = 0.0f;
                    if (row_ok && j < jmax) {
                        const bool kept = (kbits >> t) & 1u;
                        const float pr = __expf(__uint_as_float(rs[t]) + (has_bias ? br0[j] : 0.0f) - lse);
                        ds = pr * ((kept ? __uint_as_float(rp[t]) * dkey.inv_keep : 0.0f) - Di);
                    }
                    dsv[t] = ds;
                }
                if (r < lq_pad) st_tile16(sdS + (j0 >> 6) * L.tq, r, (j0 & 63) >> 4, dsv);
                if (has_bias) diag_accumulate16(drel, dsv, j0 - (warp & 3) * 32 + (TILE - 1), lane);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();                                   // every thread is done with S / dP in TMEM: the gradients may overwrite them
        tc_fence_after();
        if (tid == 0) {
            const uint32_t pa = smem_u32(sP), dsa = smem_u32(sdS);
            const uint32_t qa = smem_u32(sQ), doa = smem_u32(sdO), ka = smem_u32(sK);
            // dV = P~^T dO ; dK = dS^T Q : M = 128 keys (MN-major A: the two 64-key blocks are tq bytes apart), contraction over lq_pad queries
            for (int ks = 0; ks < nqs; ++ks) {
                umma_bf16(tmem + TM_DV, umma_smem_desc_sw128(pa + ks * 2048, L.tq, 1024), umma_smem_desc_sw128(doa + ks * 2048, 8192, 1024), id_t, ks != 0);
                umma_bf16(tmem + TM_DK, umma_smem_desc_sw128(dsa + ks * 2048, L.tq, 1024), umma_smem_desc_sw128(qa + ks * 2048, 8192, 1024), id_t, ks != 0);
            }
            // dQ = dS K : A K-major (64-key blocks tq bytes apart), B = K MN-major, contraction over lk_pad keys
            for (int ks = 0; ks < nks; ++ks)
                umma_bf16(tmem + TM_DQ, umma_smem_desc_sw128(dsa + (ks >> 2) * L.tq + (ks & 3) * 32, 16, 1024),
                          umma_smem_desc_sw128(ka + ks * 2048, 8192, 1024), id_q, ks != 0);
            umma_commit(&bars[1]);
        }
        // bias gradient, overlapped with the MMAs: diagonal sums (formed in pass 2, see diag_accumulate16) -> buckets
        if (has_bias && tid < 2 * TILE - 1) {
            const int rr = tid - (TILE - 1) + (Lq - 1);       // relative position j - i + (Lq - 1) of diagonal `tid`
            const float acc = drel[tid];
            if (rr >= 0 && rr < Lq + Lk - 1 && acc != 0.0f) atomicAdd(&bins[a.rel_bucket[rr - (Lq - 1) - a.q_offset + a.rel_zero]], acc);
        }
        mbar_wait(&bars[1], mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        if (tid == 0 && n1 < nprob) issue_loads(n1);       // the gradient MMAs have retired: all four operand tiles are free
        // write-out (row = TMEM lane): key half 0 stores columns 0..31 of dq, dk, dv, half 1 columns 32..63
#pragma unroll 1
        for (int which = 0; which < 3; ++which) {
            const int rows = which == 0 ? Lq : Lk;
            const int rows_ok = G > 1 ? min(TILE, (a.B - b) * rows) : rows;      // valid rows of the tile
            const uint32_t col = which == 0 ? TM_DQ : (which == 1 ? TM_DK : TM_DV);
            __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(which == 0 ? a.dq : (which == 1 ? a.dk : a.dv));
            const long long ld = which == 0 ? a.ldq : (which == 1 ? a.ldk : a.ldv);
            uint32_t ro[32];
            tmem_ld_32x32(trow + col + ch * 32, ro);
            tmem_ld_wait();
            if (r < rows_ok) {
                uint4* dst = reinterpret_cast<uint4*>(base + (static_cast<long long>(b) * rows + r) * ld + h * DK + ch * 32);
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    uint4 q;
                    q.x = pack_bf16(__uint_as_float(ro[8 * gq]), __uint_as_float(ro[8 * gq + 1]));
                    q.y = pack_bf16(__uint_as_float(ro[8 * gq + 2]), __uint_as_float(ro[8 * gq + 3]));
                    q.z = pack_bf16(__uint_as_float(ro[8 * gq + 4]), __uint_as_float(ro[8 * gq + 5]));
                    q.w = pack_bf16(__uint_as_float(ro[8 * gq + 6]), __uint_as_float(ro[8 * gq + 7]));
                    dst[gq] = q;
                }
            }
        }
        tc_fence_before();
        __syncthreads();                                   // TMEM, sP / sdS, brel / bins are reused by the next problem
        tc_fence_after();
        // partial bias gradient of the tile: into the slot of its first (batch, head); the other packed problems' slots get zeros
        if (has_bias && tid < a.num_buckets) {
            for (int g = 0; g < G && b + g < a.B; ++g)
                a.dbias_partial[(static_cast<long long>(b + g) * a.H + h) * a.num_buckets + tid] = g == 0 ? bins[tid] : 0.0f;
        }
        bh = *s_next;
        n1 = n2;
    }
    if (tid == 0) wc.finish_end();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

bool two_cta_enabled() {
    static const bool on = []() { const char* e = getenv("KLAB_T5_ATTN_2CTA"); return !(e && e[0] == '0'); }();
    return on;
}

// dtable[bucket, h] += sum_b partial[(b*H + h), bucket]: one WARP per (bucket, h), lanes stride over the batch, shuffle reduction
// (fixed order: deterministic).  One thread per output with a serial loop over B = 64 cost 12.5 us per launch -- pure dependent-load
// latency on the backward critical path, 48 launches per step.
__global__ void __launch_bounds__(256) t5_dbias_reduce_tc_kernel(const float* __restrict__ part, int B, int H, int nb, float* __restrict__ dtable) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nb * H) return;                                 // (warp-uniform)
    const int bucket = w / H, h = w % H;
    float s = 0.0f;
    for (int b = lane; b < B; b += 32) s += part[(static_cast<long long>(b) * H + h) * nb + bucket];
    s = warp_sum(s);
    if (lane == 0) dtable[w] += s;
}

// Problems per 128-row tile of the single-tile kernels: self-attention shapes whose length divides the tile (the 32 keys a
// thread owns must not straddle two problems, hence 32 or 64).  KLAB_T5_ATTN_PACK=0 switches packing off.
int pack_factor(int Lq, int Lk, int q_offset) {
    static const bool off = []() { const char* e = getenv("KLAB_T5_ATTN_PACK"); return e && e[0] == '0'; }();
    if (off || Lq != Lk || q_offset != 0 || (Lq != 32 && Lq != 64)) return 1;
    return TILE / Lq;
}

int make_head_map(CUtensorMap* m, const void* base, long long rows, int H, long long ld, int box_rows) {
    return make_tmap_2d_bf16(m, base, static_cast<uint64_t>(rows), static_cast<uint64_t>(H) * DK, static_cast<uint64_t>(ld), box_rows, DK);
}

}  // namespace

bool t5_attention_tc_supported(int dtype, int Lq, int Lk, int d_kv, long long ldq, long long ldk, long long ldv, long long ldo,
                               const void* q, const void* k, const void* v, const void* o) {
    if (dtype != KLAB_BF16 || d_kv != DK || Lq > 256 || Lk > 256) return false;
    if ((ldq | ldk | ldv | ldo) % 8) return false;
    return ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(o)) & 15) == 0;
}

int t5_attention_fwd_tc(cudaStream_t st, int B, int H, int Lq, int Lk, const void* q, long long ldq, const void* k, long long ldk,
                        const void* v, long long ldv, void* out, long long ldo, const float* bias_table, const int* rel_bucket,
                        int rel_zero, int num_buckets, int causal, int q_offset, float* lse, float dropout_p, unsigned long long seed,
                        const unsigned long long* seed_ptr) {
    if (Lq <= TILE && Lk <= TILE && !getenv("KLAB_T5_ATTN_MULTI")) {          // persistent single-tile kernel
        const int pack = pack_factor(Lq, Lk, q_offset);
        const int lkp = pack > 1 ? TILE : (Lk + 15) / 16 * 16;
        CUtensorMap tq, tk, tv;
        if (int rc = make_head_map(&tq, q, 1ll * B * Lq, H, ldq, TILE)) return rc;
        if (int rc = make_head_map(&tk, k, 1ll * B * Lk, H, ldk, TILE)) return rc;
        if (int rc = make_head_map(&tv, v, 1ll * B * Lk, H, ldv, TILE)) return rc;
        TcArgs a{};
        a.out = out; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk;
        a.bias_table = bias_table; a.rel_bucket = rel_bucket; a.rel_zero = rel_zero; a.num_buckets = num_buckets; a.causal = causal;
        a.q_offset = q_offset; a.lse = lse; a.dropout_p = dropout_p; a.seed = seed; a.seed_ptr = seed_ptr;
        const size_t smem1 = 1024 + 8 * TILE * 128 + sizeof(float) * (256 + 8 * TILE) + 64;
        static bool set1 = false;
        if (!set1) {
            KLAB_CHECK_CUDA(cudaFuncSetAttribute(t5_attn_fwd_tc1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            set1 = true;
        }
        a.pack = pack;
        const int nprob = ((B + pack - 1) / pack) * H;
        a.sched = sched_slot(st);
        const int sms = a.sched ? sm_count_physical() : sm_count();
        t5_attn_fwd_tc1_kernel<<<nprob < sms ? nprob : sms, ST_THREADS, smem1, st>>>(tq, tk, tv, a, lkp);
        KLAB_LAUNCH_CHECK();
        count_launch();
        return KLAB_OK;
    }
    const int lk_pad = (Lk + 15) / 16 * 16;
    const int o_col = lk_pad <= 64 ? 64 : (lk_pad <= 192 ? 192 : 256);
    const int tmem_cols = o_col + 64 <= 128 ? 128 : (o_col + 64 <= 256 ? 256 : 512);
    CUtensorMap tq, tk, tv;
    if (int rc = make_head_map(&tq, q, 1ll * B * Lq, H, ldq, TILE)) return rc;
    if (int rc = make_head_map(&tk, k, 1ll * B * Lk, H, ldk, lk_pad)) return rc;
    if (int rc = make_head_map(&tv, v, 1ll * B * Lk, H, ldv, lk_pad)) return rc;
    TcArgs a{};
    a.out = out; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk;
    a.bias_table = bias_table; a.rel_bucket = rel_bucket; a.rel_zero = rel_zero; a.num_buckets = num_buckets; a.causal = causal;
    a.q_offset = q_offset; a.lse = lse; a.dropout_p = dropout_p; a.seed = seed; a.seed_ptr = seed_ptr;
    const int nkb = (lk_pad + 63) / 64;
    const size_t smem = 1024 + TILE * 128 + 2 * lk_pad * 128 + nkb * TILE * 128 + sizeof(float) * ((Lq + Lk + 3) & ~3) + 64;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        KLAB_CHECK_CUDA(cudaFuncSetAttribute(t5_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const dim3 grid((Lq + TILE - 1) / TILE, H, B);
    t5_attn_fwd_tc_kernel<<<grid, 128, smem, st>>>(tq, tk, tv, a, lk_pad, o_col, tmem_cols);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

long long t5_attention_bwd_tc_workspace_bytes(int B, int H, int num_buckets) { return sizeof(float) * 1ll * B * H * num_buckets; }

int t5_attention_bwd_tc(cudaStream_t st, int B, int H, int Lq, int Lk, const void* q, long long ldq, const void* k, long long ldk,
                        const void* v, long long ldv, const void* out, const void* dout, long long ldo, void* dq, void* dk, void* dv,
                        const float* bias_table, const int* rel_bucket, int rel_zero, int num_buckets, int causal, int q_offset,
                        const float* lse, float* dbias_table, float dropout_p, unsigned long long seed,
                        const unsigned long long* seed_ptr, void* workspace) {
    CUtensorMap tq, tk, tv, tdo;
    if (int rc = make_head_map(&tq, q, 1ll * B * Lq, H, ldq, TILE)) return rc;
    if (int rc = make_head_map(&tk, k, 1ll * B * Lk, H, ldk, TILE)) return rc;
    if (int rc = make_head_map(&tv, v, 1ll * B * Lk, H, ldv, TILE)) return rc;
    if (int rc = make_head_map(&tdo, dout, 1ll * B * Lq, H, ldo, TILE)) return rc;
    TcArgs a{};
    a.o = out; a.dout = dout; a.dq = dq; a.dk = dk; a.dv = dv;
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk;
    a.bias_table = bias_table; a.rel_bucket = rel_bucket; a.rel_zero = rel_zero; a.num_buckets = num_buckets; a.causal = causal;
    a.q_offset = q_offset; a.lse = const_cast<float*>(lse); a.dropout_p = dropout_p; a.seed = seed; a.seed_ptr = seed_ptr;
    a.dbias_partial = static_cast<float*>(workspace);
    if (Lq <= TILE && Lk <= TILE && num_buckets <= 64 && !getenv("KLAB_T5_ATTN_MULTI") && two_cta_enabled()) {
        const int pack = pack_factor(Lq, Lk, q_offset);
        const int lqp = pack > 1 ? TILE : (Lq + 15) / 16 * 16, lkp = pack > 1 ? TILE : (Lk + 15) / 16 * 16;
        const S2Layout l2 = s2_layout(lqp, lkp);
        if (l2.total + 1024 <= S2_MAX_SMEM) {                                    // two CTAs per SM (tiles as tall as the problem)
            if (int rc = make_head_map(&tq, q, 1ll * B * Lq, H, ldq, lqp)) return rc;
            if (int rc = make_head_map(&tk, k, 1ll * B * Lk, H, ldk, lkp)) return rc;
            if (int rc = make_head_map(&tv, v, 1ll * B * Lk, H, ldv, lkp)) return rc;
            if (int rc = make_head_map(&tdo, dout, 1ll * B * Lq, H, ldo, lqp)) return rc;
            static bool set2 = false;
            if (!set2) {
                KLAB_CHECK_CUDA(cudaFuncSetAttribute(t5_attn_bwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_MAX_SMEM));
                set2 = true;
            }
            a.pack = pack;
            const int nprob = ((B + pack - 1) / pack) * H;
            a.sched = sched_slot(st);
            const int slots = 2 * (a.sched ? sm_count_physical() : sm_count());
            t5_attn_bwd_tc2_kernel<<<nprob < slots ? nprob : slots, S2_THREADS, l2.total + 1024, st>>>(tq, tk, tv, tdo, a, lqp, lkp);
            KLAB_LAUNCH_CHECK();
            count_launch();
            if (bias_table && dbias_table) {
                const int n = num_buckets * H;
                t5_dbias_reduce_tc_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(a.dbias_partial, B, H, num_buckets, dbias_table);
                KLAB_LAUNCH_CHECK();
                count_launch();
            }
            return KLAB_OK;
        }
    }
    if (Lq <= TILE && Lk <= TILE && num_buckets <= 64 && !getenv("KLAB_T5_ATTN_MULTI")) {      // persistent single-tile kernel
        const size_t smem1 = 1024 + 12 * TILE * 128 + sizeof(float) * (256 + 256 + 64 + 4 * TILE) + 64;
        static bool set1 = false;
        if (!set1) {
            KLAB_CHECK_CUDA(cudaFuncSetAttribute(t5_attn_bwd_tc1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            set1 = true;
        }
        a.pack = pack_factor(Lq, Lk, q_offset);
        const int nprob = ((B + a.pack - 1) / a.pack) * H;
        a.sched = sched_slot(st);
        const int sms = a.sched ? sm_count_physical() : sm_count();
        t5_attn_bwd_tc1_kernel<<<nprob < sms ? nprob : sms, ST_THREADS, smem1, st>>>(tq, tk, tv, tdo, a, a.pack > 1 ? TILE : (Lq + 15) / 16 * 16,
                                                                                                   a.pack > 1 ? TILE : (Lk + 15) / 16 * 16);
        KLAB_LAUNCH_CHECK();
        count_launch();
        if (bias_table && dbias_table) {
            const int n = num_buckets * H;
            t5_dbias_reduce_tc_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(a.dbias_partial, B, H, num_buckets, dbias_table);
            KLAB_LAUNCH_CHECK();
            count_launch();
        }
        return KLAB_OK;
    }
    const int nq = (Lq + TILE - 1) / TILE, nk = (Lk + TILE - 1) / TILE;
    const size_t smem = 1024 + static_cast<size_t>(2 * nq + 2 * nk + 4) * TILE * 128 + 2 * sizeof(float) * ((Lq + Lk + 3) & ~3) +
                        sizeof(float) * ((num_buckets + 3) & ~3) + 64;
    KLAB_REQUIRE(smem <= 227 * 1024, "t5_attention_bwd_tc: %zu bytes of shared memory", smem);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        KLAB_CHECK_CUDA(cudaFuncSetAttribute(t5_attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const dim3 grid(H, B);
    t5_attn_bwd_tc_kernel<<<grid, 128, smem, st>>>(tq, tk, tv, tdo, a);
    KLAB_LAUNCH_CHECK();
    count_launch();
    if (bias_table && dbias_table) {
        const int n = num_buckets * H;
        t5_dbias_reduce_tc_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(a.dbias_partial, B, H, num_buckets, dbias_table);
        KLAB_LAUNCH_CHECK();
        count_launch();
    }
    return KLAB_OK;
}

}  // namespace klab
