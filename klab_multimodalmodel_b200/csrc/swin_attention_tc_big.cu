// K2 + K3 + K4 for LARGE windows: Swin-V2 shifted-window cosine attention on tcgen05 for 65..160 tokens per window, head_dim 32
// (the 12 x 12 = 144-token windows of 384^2 inputs, BASELINE config 4; also 9x9 .. 12x12).  Semantics identical to
// swin_attention.cu / swin_attention_tc.cu (HF/models/swinv2/modeling_swinv2.py:421-487, window partition / cyclic roll / shift
// mask folded into index math, -200 mask as the reference adds it twice).
//
// A window no longer fits the two-windows-per-tile scheme of swin_attention_tc.cu, so a work item is ONE (head, window):
//   * all N keys of the window are staged once (L2-normalised k as bf16 hi | lo, v) as [160 x 64] K-major / MN-major tiles;
//   * the query rows go through the 128-row UMMA tile in PASSES: pass 0 = tokens 0..127, pass 1 = tokens 128..N-1 (the rest of
//     the tile is padding).  S = Qh Kh^T (N_mma = 144) lands in TMEM, the softmax runs one thread per row straight out of
//     TMEM, P goes back through shared memory as the A operand of O = P V;
//   * backward keeps S and dP ([128 x 144] each) in TMEM, accumulates dK and dV of the window over both passes in TMEM
//     (M = keys: two M-tiles, 0..127 and 128..N-1), writes dQ per pass; the position-bias gradient of the rows of pass 0 lives
//     in registers across all windows a CTA visits for a head, that of the few rows of pass 1 in shared memory.
// Position bias: a [144 x 144] fp32 table per head does not fit next to the operand tiles, and reading a query's bias row from
// L2 inside the softmax loop exposes the L2 latency once per 16 keys.  The bias is a function of the relative offset only
// (bias[i][j] = tab[(yi - yj + w - 1)(2w - 1) + (xi - xj + w - 1)], HF modeling_swinv2.py:512-522), so the BACKWARD kernel rebuilds the
// (2w - 1)^2-entry table of the current head in shared memory (2 KB) from the gathered bias and indexes it as
// tab[rowbase(i) - joff[j]] with a per-key offset table (the forward kernel keeps the row reads: see load_bias16).
#include "swin_tc.cuh"

namespace klab {
void count_launch(int n = 1);
int sm_count();
namespace {
using namespace swintc;

constexpr int KROWS = 160;                     // staged key rows (rows >= N are zero); 20 groups of 8 rows
constexpr int KBYTES = KROWS * 128;            // bytes of the K / V tiles
constexpr int NMAX = 160;                      // TMEM budget of the backward kernel: S + dP = 2 * 160 columns
constexpr int TABMAX = 640;                    // >= (2w - 1)^2 = 625 entries of the relative-offset bias table (w <= 13); even: the mbarriers behind it stay 8-byte aligned

// relative-offset table of one head from the gathered [N, N] bias: entry d = (dy + w - 1)(2w - 1) + (dx + w - 1) is bias[i][j] of any
// pair with (yi - yj, xi - xj) = (dy, dx); pick i = (max(dy, 0), max(dx, 0)), j = (max(-dy, 0), max(-dx, 0))
__device__ __forceinline__ void build_bias_table(float* tab_s, const float* bias_h, int w, int N, int tid, int nthreads) {
    const int side = 2 * w - 1;
    for (int d = tid; d < side * side; d += nthreads) {
        const int dy = d / side - (w - 1), dx = d % side - (w - 1);
        const int i = max(dy, 0) * w + max(dx, 0), j = max(-dy, 0) * w + max(-dx, 0);
        tab_s[d] = __ldg(bias_h + static_cast<long long>(i) * N + j);
    }
}
__device__ __forceinline__ int bias_rowbase(int n, int w) { return (n / w + w - 1) * (2 * w - 1) + (n % w + w - 1); }

// 16 consecutive entries of a query's bias row straight from global memory (forward kernel: the row is 576 contiguous bytes per
// thread, read twice and L1-resident the second time -- measured 14 % faster there than the shared-memory table, whose two
// dependent LDS per key sit on the softmax critical path; the backward kernel, 4 threads per row, is 12 % faster WITH the table)
__device__ __forceinline__ void load_bias16(const float* brow, int c0, int N, bool vec, float* b) {
    if (vec && c0 + 16 <= N) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(brow + c0) + i);
            b[4 * i] = q.x; b[4 * i + 1] = q.y; b[4 * i + 2] = q.z; b[4 * i + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] = c0 + i < N ? __ldg(brow + c0 + i) : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// forward: 128 threads, thread t = query row t of the current pass; two CTAs per SM (112 KB of shared memory, 256 TMEM columns
// each).  Work item = (head, window, pass), head-major; a CTA owns a contiguous range of items, so the second pass of a
// window finds the window's keys / values still staged.
// TMEM: S 0..nk-1 | O 192..223
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) swin_attn_fwd_big_kernel(SwinTcArgs a, int passes, int nk) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + smem_align_pad(smem_raw);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TB;
    uint8_t* sV = sK + KBYTES;
    uint8_t* sP = sV + KBYTES;                               // 3 key blocks of 64
    int* sregk = reinterpret_cast<int*>(sP + 3 * TB);        // [KROWS] shift-mask region of every key token
    uint64_t* bars = reinterpret_cast<uint64_t*>(sregk + KROWS);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);

    const int N = a.N;
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool vec = (N & 3) == 0;
    const int nwin = a.B * a.nW;
    const long long T = static_cast<long long>(a.heads) * nwin * passes;
    const long long t_begin = blockIdx.x * T / gridDim.x, t_end = (blockIdx.x + 1) * T / gridDim.x;
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
    constexpr int O_COL = 192;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t id_s = umma_idesc_bf16(TILE, nk, false, false), id_o = umma_idesc_bf16(TILE, HD, false, true);
    uint32_t phase = 0;
    int h = -1;
    long long staged = -1;                                   // (head, window) whose keys / values are in sK / sV
    float scale = 0.0f;

    for (long long t = t_begin; t < t_end; ++t) {
        const int th = static_cast<int>(t / (static_cast<long long>(nwin) * passes));
        const int rem = static_cast<int>(t - static_cast<long long>(th) * nwin * passes);
        const int bw = rem / passes, pass = rem - bw * passes;
        if (th != h) {
            h = th;
            scale = __expf(fminf(a.logit_scale[h], LOGIT_MAX));
        }
        // ---- stage this pass's query row, and the window's keys / values if they are not there yet ----
        const int n = pass * TILE + tid;
        int region;
        const int tok = window_token(a, bw, n, region);
        {
            float q[HD];
            if (tok >= 0) {
                load_row32(a.q, a.ld, tok, h, q);
                float nq;
                const float iq = inv_norm32(q, nq);
#pragma unroll
                for (int c = 0; c < HD; ++c) q[c] *= iq;
            } else {
#pragma unroll
                for (int c = 0; c < HD; ++c) q[c] = 0.0f;
            }
            stage_row32_hilo(sQ, tid, q);
        }
        const long long key = static_cast<long long>(h) * nwin + bw;
        if (key != staged) {
            for (int kr = tid; kr < KROWS; kr += 128) {
                int kreg;
                const int ktok = window_token(a, bw, kr, kreg);
                float k[HD];
                if (ktok >= 0) {
                    load_row32(a.k, a.ld, ktok, h, k);
                    float nkk;
                    const float ik = inv_norm32(k, nkk);
#pragma unroll
                    for (int c = 0; c < HD; ++c) k[c] *= ik;
                    const uint4* vp = reinterpret_cast<const uint4*>(a.v + static_cast<long long>(ktok) * a.ld + h * HD);
#pragma unroll
                    for (int c = 0; c < 4; ++c) st_tile8_raw(sV, kr, c, __ldg(vp + c));
                } else {
#pragma unroll
                    for (int c = 0; c < HD; ++c) k[c] = 0.0f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) st_tile8_raw(sV, kr, c, make_uint4(0, 0, 0, 0));
                }
                stage_row32_hilo(sK, kr, k);
                sregk[kr] = kreg;
            }
            staged = key;
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            issue_cosine_logits(tmem, smem_u32(sQ), smem_u32(sK), id_s);
            umma_commit(&bars[0]);
        }
        mbar_wait(&bars[0], phase);
        tc_fence_after();

        // ---- softmax of row tid over the window's N keys, straight out of TMEM: pass A finds the maximum, pass B writes P ----
        const float* brow = a.bias + (static_cast<long long>(h) * N + (tok >= 0 ? n : 0)) * N;
        float mx = -INFINITY;
        for (int c0 = 0; c0 < nk; c0 += 16) {
            uint32_t r[16];
            tmem_ld_32x16(trow + c0, r);
            tmem_ld_wait();
            if (tok >= 0) {
                float b[16];
                load_bias16(brow, c0, N, vec, b);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int j = c0 + i;
                    if (j < N) {
                        float sc = fmaf(__uint_as_float(r[i]), scale, b[i]);
                        if (sregk[j] != region) sc += -200.0f;
                        mx = fmaxf(mx, sc);
                    }
                }
            }
        }
        float sum = 0.0f;
        for (int c0 = 0; c0 < nk; c0 += 16) {
            uint32_t r[16];
            tmem_ld_32x16(trow + c0, r);
            tmem_ld_wait();
            float e[16];
            if (tok >= 0) {
                float b[16];
                load_bias16(brow, c0, N, vec, b);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int j = c0 + i;
                    float v = 0.0f;
                    if (j < N) {
                        float sc = fmaf(__uint_as_float(r[i]), scale, b[i]);
                        if (sregk[j] != region) sc += -200.0f;
                        v = __expf(sc - mx);
                    }
                    e[i] = v;
                    sum += v;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) e[i] = 0.0f;
            }
            uint8_t* blk = sP + (c0 >> 6) * TB;
            st_tile8(blk, tid, (c0 & 63) >> 3, e);
            st_tile8(blk, tid, ((c0 & 63) >> 3) + 1, e + 8);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
            for (int ks = 0; ks < nk / 16; ++ks)
                umma_bf16(tmem + O_COL, umma_smem_desc_sw128(pa + (ks >> 2) * TB + (ks & 3) * 32, 16, 1024),
                          umma_smem_desc_sw128(va + ks * 2048, 8192, 1024), id_o, ks != 0);
            umma_commit(&bars[1]);
        }
        mbar_wait(&bars[1], phase);
        phase ^= 1;
        tc_fence_after();
        {
            uint32_t r[32];
            tmem_ld_32x32(trow + O_COL, r);
            tmem_ld_wait();
            if (tok >= 0) {
                const float inv = 1.0f / sum;
                float o[HD];
#pragma unroll
                for (int c = 0; c < HD; ++c) o[c] = __uint_as_float(r[c]) * inv;
                store_row32(a.out, a.ldc, tok, h, o);
                a.lse[(static_cast<long long>(bw) * a.heads + h) * N + n] = mx + __logf(sum);
            }
        }
        tc_fence_before();          // the next item's staging barrier orders these TMEM reads before its MMAs
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward: 512 threads, one CTA per SM; work item = (head, window), both query passes inside the item.
// TMEM: S 0.. | dP 160.. | dV keys 0..127: 320, keys 128..: 352 | dK: 384, 416 | dQ 448      (N <= 160)
// Thread view in the softmax backward: TMEM lane r = (warp & 3) * 32 + lane is the query row, cq = warp >> 2 selects 48 of the keys
// (three 16-column TMEM loads; the last group is idle for N <= 144).
// ------------------------------------------------------------------------------------------------------------------
constexpr int BIG_BWD_THREADS = 512;
constexpr int KQ = 48;

__global__ void __launch_bounds__(BIG_BWD_THREADS, 1) swin_attn_bwd_big_kernel(SwinTcArgs a, int passes, int nk) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + smem_align_pad(smem_raw);
    uint8_t* sQ = smem;
    uint8_t* sdO = sQ + TB;
    uint8_t* sK = sdO + TB;
    uint8_t* sV = sK + KBYTES;
    uint8_t* sP = sV + KBYTES;                                     // 4 key blocks (block 3 stays zero: upper half of the second M-tile)
    uint8_t* sdS = sP + 4 * TB;
    const int N = a.N;
    const int nrem = N > TILE ? N - TILE : 0;                      // query rows of pass 1
    float* dbrem = reinterpret_cast<float*>(sdS + 4 * TB);         // [nrem][N + 1] bias gradient of the rows of pass 1
    int* sreg = reinterpret_cast<int*>(dbrem + nrem * (N + 1));    // [128] region of the pass's query rows, -1 = padding
    int* sregk = sreg + TILE;                                      // [KROWS]
    int* joff = sregk + KROWS;                                     // [KROWS] yj (2w - 1) + xj of key token j
    float* tab_s = reinterpret_cast<float*>(joff + KROWS);         // [(2w - 1)^2] relative-offset bias table of the current head
    float* sqn = tab_s + TABMAX;                                   // [128] |q|
    float* skn = sqn + TILE;                                       // [KROWS] |k|
    float* sD = skn + KROWS;                                       // [4][128]
    float* red = sD + 4 * TILE;                                    // [16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 16);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
    constexpr int TM_S = 0, TM_DP = 160, TM_DV = 320, TM_DK = 384, TM_DQ = 448;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nwin = a.B * a.nW;
    const long long T = static_cast<long long>(a.heads) * nwin;
    const long long t_begin = blockIdx.x * T / gridDim.x, t_end = (blockIdx.x + 1) * T / gridDim.x;
    const int srow = tid >> 2, part = tid & 3;                     // staging view: 4 threads per token row, 16 bytes each
    const int r = (warp & 3) * 32 + lane, cq = warp >> 2;          // TMEM view
    for (int j = tid; j < KROWS; j += BIG_BWD_THREADS) joff[j] = (j / a.w) * (2 * a.w - 1) + j % a.w;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    for (int idx = tid; idx < 8 * TB / 16; idx += BIG_BWD_THREADS) reinterpret_cast<uint4*>(sP)[idx] = make_uint4(0, 0, 0, 0);
    for (int idx = tid; idx < nrem * (N + 1); idx += BIG_BWD_THREADS) dbrem[idx] = 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t id_s = umma_idesc_bf16(TILE, nk, false, false);     // S / dP
    const uint32_t id_t = umma_idesc_bf16(TILE, HD, true, true);       // dV / dK : A = P^T / dS^T (MN-major), B MN-major
    const uint32_t id_q = umma_idesc_bf16(TILE, HD, false, true);      // dQ      : A = dS (K-major), B = Kh MN-major
    int h = -1;
    float raw_ls = 0.0f, scale = 0.0f, dscale_acc = 0.0f;
    float dbacc[KQ];
    uint32_t phase = 0;

    auto flush_head = [&]() {
        if (r < N) {
            float* drow = a.dbias + (static_cast<long long>(h) * N + r) * N;
#pragma unroll
            for (int i = 0; i < KQ; ++i) {
                const int j = cq * KQ + i;
                if (j < N && dbacc[i] != 0.0f) atomicAdd(&drow[j], dbacc[i]);
            }
        }
        __syncthreads();
        for (int idx = tid; idx < nrem * N; idx += BIG_BWD_THREADS) {
            const int i = idx / N, j = idx - i * N;
            const float v = dbrem[i * (N + 1) + j];
            if (v != 0.0f) atomicAdd(&a.dbias[(static_cast<long long>(h) * N + TILE + i) * N + j], v);
            dbrem[i * (N + 1) + j] = 0.0f;
        }
        dscale_acc = warp_sum(dscale_acc);
        if (lane == 0) red[warp] = dscale_acc;
        __syncthreads();
        if (tid == 0) {
            float s = 0.0f;
#pragma unroll
            for (int w = 0; w < BIG_BWD_THREADS / 32; ++w) s += red[w];
            atomicAdd(&a.dlogit_scale[h], raw_ls <= LOGIT_MAX ? s * scale : 0.0f);
        }
        __syncthreads();
    };

    for (long long t = t_begin; t < t_end; ++t) {
        const int th = static_cast<int>(t / nwin);
        const int bw = static_cast<int>(t - static_cast<long long>(th) * nwin);
        if (th != h) {
            if (h >= 0) flush_head();
            h = th;
            raw_ls = a.logit_scale[h];
            scale = __expf(fminf(raw_ls, LOGIT_MAX));
            dscale_acc = 0.0f;
#pragma unroll
            for (int i = 0; i < KQ; ++i) dbacc[i] = 0.0f;
            build_bias_table(tab_s, a.bias + static_cast<long long>(h) * N * N, a.w, N, tid, BIG_BWD_THREADS);     // published by the staging barrier
        }
        // ---- stage the window's keys / values: L2-normalised k as bf16 hi | lo, v, |k|, region ----
        for (int idx = tid; idx < KROWS * 4; idx += BIG_BWD_THREADS) {
            const int kr = idx >> 2;
            int kreg;
            const int ktok = window_token(a, bw, kr, kreg);
            uint4 pk = make_uint4(0, 0, 0, 0), pv = pk;
            if (ktok >= 0) {
                const long long o = static_cast<long long>(ktok) * a.ld + h * HD;
                pk = __ldg(reinterpret_cast<const uint4*>(a.k + o) + part);
                pv = __ldg(reinterpret_cast<const uint4*>(a.v + o) + part);
            }
            float k[8], kl[8];
            unpack8(pk, k);
            float sk = 0.0f;
#pragma unroll
            for (int c = 0; c < 8; ++c) sk = fmaf(k[c], k[c], sk);
            sk += __shfl_xor_sync(0xffffffffu, sk, 1);
            sk += __shfl_xor_sync(0xffffffffu, sk, 2);
            const float kn = fmaxf(sqrtf(sk), NORM_EPS);
            const float ik = ktok >= 0 ? 1.0f / kn : 0.0f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                k[c] *= ik;
                kl[c] = k[c] - __bfloat162float(__float2bfloat16_rn(k[c]));
            }
            st_tile8(sK, kr, part, k);
            st_tile8(sK, kr, 4 + part, kl);
            st_tile8_raw(sV, kr, part, pv);
            if (part == 0) {
                sregk[kr] = ktok >= 0 ? kreg : -1;
                skn[kr] = kn;
            }
        }
        for (int pass = 0; pass < passes; ++pass) {
            // ---- stage this pass's query rows: normalised q (hi | lo), dO ----
            {
                const int qn_tok = pass * TILE + srow;
                int qreg;
                const int qtok = window_token(a, bw, qn_tok, qreg);
                uint4 pq = make_uint4(0, 0, 0, 0), pdo = pq;
                if (qtok >= 0) {
                    pq = __ldg(reinterpret_cast<const uint4*>(a.q + static_cast<long long>(qtok) * a.ld + h * HD) + part);
                    pdo = __ldg(reinterpret_cast<const uint4*>(a.dctx + static_cast<long long>(qtok) * a.ldc + h * HD) + part);
                }
                float q[8], ql[8];
                unpack8(pq, q);
                float sq = 0.0f;
#pragma unroll
                for (int c = 0; c < 8; ++c) sq = fmaf(q[c], q[c], sq);
                sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                const float qn = fmaxf(sqrtf(sq), NORM_EPS);
                const float iq = qtok >= 0 ? 1.0f / qn : 0.0f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    q[c] *= iq;
                    ql[c] = q[c] - __bfloat162float(__float2bfloat16_rn(q[c]));
                }
                st_tile8(sQ, srow, part, q);
                st_tile8(sQ, srow, 4 + part, ql);
                st_tile8_raw(sdO, srow, part, pdo);
                if (part == 0) {
                    sreg[srow] = qtok >= 0 ? qreg : -1;
                    sqn[srow] = qn;
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            if (tid == 0) {
                const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK), va = smem_u32(sV), doa = smem_u32(sdO);
                issue_cosine_logits(tmem + TM_S, qa, ka, id_s);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_bf16(tmem + TM_DP, umma_smem_desc_sw128(doa + k * 32, 16, 1024), umma_smem_desc_sw128(va + k * 32, 16, 1024), id_s, k != 0);
                umma_commit(&bars[0]);
            }
            const int n = pass * TILE + r;                          // this thread's query token
            const int region = sreg[r];
            const bool valid = region >= 0;
            const float lse = valid ? a.lse[(static_cast<long long>(bw) * a.heads + h) * N + n] : 0.0f;
            const int rb = bias_rowbase(valid ? n : 0, a.w);
            mbar_wait(&bars[0], phase);
            phase ^= 1;
            tc_fence_after();
            // phase 1: D_i = sum_j P_ij dP_ij over this thread's keys (P and dP recomputed in phase 2: 48 keys do not fit in registers twice)
            float Dp = 0.0f;
#pragma unroll
            for (int ci = 0; ci < KQ / 16; ++ci) {
                const int c0 = cq * KQ + ci * 16;
                if (c0 < nk) {
                    uint32_t rs[16], rp[16];
                    tmem_ld_32x16(trow + TM_S + c0, rs);
                    tmem_ld_32x16(trow + TM_DP + c0, rp);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int j = c0 + i;
                            if (j < N) {
                                float sc = fmaf(__uint_as_float(rs[i]), scale, tab_s[rb - joff[j]]);
                                if (sregk[j] != region) sc += -200.0f;
                                Dp = fmaf(__expf(sc - lse), __uint_as_float(rp[i]), Dp);
                            }
                        }
                    }
                }
            }
            sD[cq * TILE + r] = Dp;
            __syncthreads();
            const float Di = (sD[r] + sD[TILE + r]) + (sD[2 * TILE + r] + sD[3 * TILE + r]);
            // phase 2: P, dS = P (dP - D); bias / logit-scale gradients; P and dS tiles for the gradient MMAs
#pragma unroll
            for (int ci = 0; ci < KQ / 16; ++ci) {
                const int c0 = cq * KQ + ci * 16;
                if (c0 < nk) {
                    uint32_t rs[16], rp[16];
                    tmem_ld_32x16(trow + TM_S + c0, rs);
                    tmem_ld_32x16(trow + TM_DP + c0, rp);
                    tmem_ld_wait();
                    float pv[16], dsv[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) { pv[i] = 0.0f; dsv[i] = 0.0f; }
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int j = c0 + i;
                            if (j < N) {
                                float sc = fmaf(__uint_as_float(rs[i]), scale, tab_s[rb - joff[j]]);
                                if (sregk[j] != region) sc += -200.0f;
                                const float p = __expf(sc - lse);
                                const float ds = p * (__uint_as_float(rp[i]) - Di);
                                pv[i] = p;
                                dsv[i] = ds;
                                dscale_acc = fmaf(ds, __uint_as_float(rs[i]), dscale_acc);
                                if (pass == 0) dbacc[ci * 16 + i] += ds;
                                else dbrem[r * (N + 1) + j] += ds;
                            }
                        }
                    }
                    uint8_t* pb = sP + (c0 >> 6) * TB;
                    uint8_t* db = sdS + (c0 >> 6) * TB;
                    const int c8 = (c0 & 63) >> 3;
                    st_tile8(pb, r, c8, pv);
                    st_tile8(pb, r, c8 + 1, pv + 8);
                    st_tile8(db, r, c8, dsv);
                    st_tile8(db, r, c8 + 1, dsv + 8);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            if (tid == 0) {
                const uint32_t pa = smem_u32(sP), dsa = smem_u32(sdS);
                const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK), doa = smem_u32(sdO);
                // dV (+)= P^T dO, dK (+)= dS^T Qh over the query rows of this pass (K = 128 rows); two M-tiles of keys
#pragma unroll
                for (int ks = 0; ks < TILE / 16; ++ks) {
                    const uint32_t acc = (pass | ks) != 0;
                    const uint64_t bdo = umma_smem_desc_sw128(doa + ks * 2048, 8192, 1024), bq = umma_smem_desc_sw128(qa + ks * 2048, 8192, 1024);
                    umma_bf16(tmem + TM_DV, umma_smem_desc_sw128(pa + ks * 2048, TB, 1024), bdo, id_t, acc);
                    umma_bf16(tmem + TM_DV + HD, umma_smem_desc_sw128(pa + 2 * TB + ks * 2048, TB, 1024), bdo, id_t, acc);
                    umma_bf16(tmem + TM_DK, umma_smem_desc_sw128(dsa + ks * 2048, TB, 1024), bq, id_t, acc);
                    umma_bf16(tmem + TM_DK + HD, umma_smem_desc_sw128(dsa + 2 * TB + ks * 2048, TB, 1024), bq, id_t, acc);
                }
                // dQ = dS Kh over the keys
                for (int kk = 0; kk < nk / 16; ++kk)
                    umma_bf16(tmem + TM_DQ, umma_smem_desc_sw128(dsa + (kk >> 2) * TB + (kk & 3) * 32, 16, 1024),
                              umma_smem_desc_sw128(ka + kk * 2048, 8192, 1024), id_q, kk != 0);
                umma_commit(&bars[0]);
            }
            mbar_wait(&bars[0], phase);
            phase ^= 1;
            tc_fence_after();
            // ---- dq of this pass (warps 0..3): d q = (d qh - qh (qh . d qh)) / |q| with d qh = scale * (dS Kh) ----
            if (cq == 0) {
                uint32_t acc[32];
                tmem_ld_32x32(trow + TM_DQ, acc);
                tmem_ld_wait();
                int rg;
                const int tok = window_token(a, bw, n, rg);
                if (tok >= 0) {
                    float xh[HD], o[HD];
#pragma unroll
                    for (int c = 0; c < 4; ++c) unpack8(ld_tile8_raw(sQ, r, c), xh + 8 * c);
                    float dot = 0.0f;
#pragma unroll
                    for (int c = 0; c < HD; ++c) {
                        o[c] = __uint_as_float(acc[c]) * scale;
                        dot = fmaf(o[c], xh[c], dot);
                    }
                    const float inv = 1.0f / sqn[r];
#pragma unroll
                    for (int c = 0; c < HD; ++c) o[c] = (o[c] - xh[c] * dot) * inv;
                    store_row32(a.dq, a.ld, tok, h, o);
                }
            }
            tc_fence_before();
            __syncthreads();          // sQ / sdO / sP / sdS and the S / dP / dQ accumulators are reused by the next pass
            tc_fence_after();
        }
        // ---- dk / dv of the window: M-tile 0 (keys 0..127) by warps 4..11, M-tile 1 (keys 128..) by warps 12..15 ----
        {
            const int mt = cq == 3 ? 1 : 0;
            const int kr = mt * TILE + r;                              // key token of this thread's TMEM lane
            const bool do_k = cq == 1 || cq == 3, do_v = cq == 2 || cq == 3;
            int rg;
            const int tok = (cq >= 1 && kr < KROWS) ? window_token(a, bw, kr, rg) : -1;
            if (do_k) {
                uint32_t acc[32];
                tmem_ld_32x32(trow + TM_DK + mt * HD, acc);
                tmem_ld_wait();
                if (tok >= 0) {
                    float xh[HD], o[HD];
#pragma unroll
                    for (int c = 0; c < 4; ++c) unpack8(ld_tile8_raw(sK, kr, c), xh + 8 * c);
                    float dot = 0.0f;
#pragma unroll
                    for (int c = 0; c < HD; ++c) {
                        o[c] = __uint_as_float(acc[c]) * scale;
                        dot = fmaf(o[c], xh[c], dot);
                    }
                    const float inv = 1.0f / skn[kr];
#pragma unroll
                    for (int c = 0; c < HD; ++c) o[c] = (o[c] - xh[c] * dot) * inv;
                    store_row32(a.dk, a.ld, tok, h, o);
                }
            }
            if (do_v) {
                uint32_t acc[32];
                tmem_ld_32x32(trow + TM_DV + mt * HD, acc);
                tmem_ld_wait();
                if (tok >= 0) {
                    float o[HD];
#pragma unroll
                    for (int c = 0; c < HD; ++c) o[c] = __uint_as_float(acc[c]);
                    store_row32(a.dv, a.ld, tok, h, o);
                }
            }
        }
        tc_fence_before();
        __syncthreads();              // sK / sV and the dK / dV accumulators are reused by the next window
        tc_fence_after();
    }
    if (h >= 0) flush_head();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

SwinTcArgs make_big_args(int B, int res, int heads, int window, int shift, const void* q, const void* k, const void* v, long long ld,
                         long long ldc, const float* logit_scale, const float* bias, float* lse) {
    SwinTcArgs a{};
    a.q = static_cast<const __nv_bfloat16*>(q); a.k = static_cast<const __nv_bfloat16*>(k); a.v = static_cast<const __nv_bfloat16*>(v);
    a.ld = ld; a.ldc = ldc; a.B = B; a.res = res; a.heads = heads; a.w = window; a.shift = shift; a.N = window * window;
    a.nW = (res / window) * (res / window);
    a.logit_scale = logit_scale; a.bias = bias; a.lse = lse;
    return a;
}

}  // namespace

bool swin_attention_big_supported(int dtype, int head_dim, int window, long long ld, long long ldc, const void* q, const void* k,
                                  const void* v, const void* ctx) {
    const int N = window * window;
    if (dtype != KLAB_BF16 || head_dim != HD || N <= 64 || N > NMAX) return false;
    if (N > TILE + 16) return false;                  // shared memory of the backward kernel: at most 16 query rows in the second pass
    if ((ld | ldc) % 8) return false;
    return ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(ctx)) & 15) == 0;
}

int swin_attention_fwd_big(cudaStream_t st, int B, int res, int heads, int window, int shift, const void* q, const void* k, const void* v,
                           long long ld, void* ctx, long long ldc, const float* logit_scale, const float* bias, float* lse) {
    SwinTcArgs a = make_big_args(B, res, heads, window, shift, q, k, v, ld, ldc, logit_scale, bias, lse);
    a.out = static_cast<__nv_bfloat16*>(ctx);
    const int N = a.N, passes = (N + TILE - 1) / TILE, nk = (N + 15) & ~15;
    const size_t smem = 1024 + TB + 2 * KBYTES + 3 * TB + sizeof(int) * KROWS + 64;
    static bool set = false;
    if (!set) {
        KLAB_CHECK_CUDA(cudaFuncSetAttribute(swin_attn_fwd_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        set = true;
    }
    const long long items = 1ll * heads * B * a.nW * passes;
    long long grid = 2ll * sm_count();                                 // two resident CTAs per SM
    if (grid > items) grid = items;
    swin_attn_fwd_big_kernel<<<static_cast<unsigned>(grid), 128, smem, st>>>(a, passes, nk);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

int swin_attention_bwd_big(cudaStream_t st, int B, int res, int heads, int window, int shift, const void* q, const void* k, const void* v,
                           long long ld, const void* ctx, const void* dctx, long long ldc, void* dq, void* dk, void* dv,
                           const float* logit_scale, const float* bias, const float* lse, float* dbias, float* dlogit_scale) {
    SwinTcArgs a = make_big_args(B, res, heads, window, shift, q, k, v, ld, ldc, logit_scale, bias, const_cast<float*>(lse));
    a.ctx = static_cast<const __nv_bfloat16*>(ctx); a.dctx = static_cast<const __nv_bfloat16*>(dctx);
    a.dq = static_cast<__nv_bfloat16*>(dq); a.dk = static_cast<__nv_bfloat16*>(dk); a.dv = static_cast<__nv_bfloat16*>(dv);
    a.dbias = dbias; a.dlogit_scale = dlogit_scale;
    const int N = a.N, passes = (N + TILE - 1) / TILE, nk = (N + 15) & ~15;
    const int nrem = N > TILE ? N - TILE : 0;
    const size_t smem = 1024 + 2 * TB + 2 * KBYTES + 8 * TB + sizeof(float) * nrem * (N + 1) + sizeof(int) * (TILE + 2 * KROWS) +
                        sizeof(float) * (TABMAX + TILE + KROWS + 4 * TILE + 16) + 64;
    KLAB_REQUIRE(smem <= 227 * 1024, "swin_attention_bwd (large windows): %zu bytes of shared memory", smem);
    static size_t set = 0;
    if (smem > set) {
        KLAB_CHECK_CUDA(cudaFuncSetAttribute(swin_attn_bwd_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        set = smem;
    }
    KLAB_CHECK_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * heads * N * N, st));
    KLAB_CHECK_CUDA(cudaMemsetAsync(dlogit_scale, 0, sizeof(float) * heads, st));
    const long long items = 1ll * heads * B * a.nW;
    long long grid = sm_count();
    if (grid > items) grid = items;
    swin_attn_bwd_big_kernel<<<static_cast<unsigned>(grid), BIG_BWD_THREADS, smem, st>>>(a, passes, nk);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

}  // namespace klab
