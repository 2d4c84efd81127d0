// K2 + K3 + K4 (hot path): Swin-V2 shifted-window cosine attention on the 5th-gen tensor cores, head_dim = 32, window <= 8x8.
//
// Same semantics as swin_attention.cu (HF/models/swinv2/modeling_swinv2.py:421-487 with window partition / cyclic roll /
// shift mask folded into index math); this file moves the arithmetic onto tcgen05:
//   * two windows share one 128-row UMMA tile (window slot g = row / 64); S = Qh Kh^T is computed for all 128 x 128 pairs and
//     the cross-window half is simply never read; P of the other slot is written as zeros so O = P V needs no masking;
//   * the token gather (roll + partition), the L2 normalisation of q and k (fp32) and the bf16 conversion happen while the
//     operands are staged into 128B-swizzled shared memory -- q/k/v/ctx stay in natural token order in HBM;
//   * S, P, dP, dS live in TMEM / shared memory only; backward keeps the bias gradient of all windows a CTA visits in shared
//     memory and flushes it once (one atomicAdd per element per CTA instead of per window).
// Limits: bf16, head_dim == 32, window*window <= 64 (7x7 and 8x8 of the BASELINE geometries); 12x12 windows (384^2 inputs)
// and fp32 take the generic kernel.
#include "swin_tc.cuh"

namespace klab {
void count_launch(int n = 1);
namespace {
using namespace swintc;

// ------------------------------------------------------------------------------------------------------------------
// forward: 1-D grid of two CTAs per SM, 128 threads; thread t = row t = (slot t/64, token t%64).  Work item = (head, window
// pair), numbered head-major; CTA c owns the contiguous range [c T / G, (c + 1) T / G) of the T = heads * npairs items, so every
// CTA gets the same number of items (+-1) whatever the head count, and changes head at most a few times.  The current head's
// position bias (identical for every window) sits in shared memory with rows padded to N + 1 floats -- lanes read different query rows, which from global memory costs
// one cache line per lane per key (it used to dominate the kernel) -- and the q / k / v rows of the next pair are prefetched
// into registers while the tensor core and the softmax work on the current one.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) swin_attn_fwd_tc_kernel(SwinTcArgs a, int npairs) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + smem_align_pad(smem_raw);      // 1024-byte aligned, still a __shared__ pointer (LDS / STS, 32-bit addressing)
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TB;
    uint8_t* sV = sK + TB;
    uint8_t* sP = sV + TB;                               // 2 key blocks of 64
    const int N = a.N;
    float* sB = reinterpret_cast<float*>(sP + 2 * TB);   // [N][N + 1] bias of this head
    int* sreg = reinterpret_cast<int*>(sB + N * (N + 1));   // [128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sreg + TILE);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int g = tid >> 6, n = tid & 63;
    const int T = npairs * a.heads;
    const int t_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * T / gridDim.x);
    const int t_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * T / gridDim.x);
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
    // the other window slot's key block of P stays zero for the whole kernel (cross-window half of the 128 x 128 product)
#pragma unroll
    for (int c = 0; c < 8; ++c) st_tile8_raw(sP + (1 - g) * TB, tid, c, make_uint4(0, 0, 0, 0));

    // prefetch registers: the q / k / v head slices of this thread's row of the pair about to be staged
    uint4 pq[4], pk[4], pv[4];
    int ptok = -1, pregion = 0;
    auto prefetch = [&](int t) {
        const int ph = t / npairs, pair = t - ph * npairs;
        ptok = window_token(a, pair * 2 + g, n, pregion);
        if (ptok >= 0) {
            const long long o = static_cast<long long>(ptok) * a.ld + ph * HD;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                pq[c] = __ldg(reinterpret_cast<const uint4*>(a.q + o) + c);
                pk[c] = __ldg(reinterpret_cast<const uint4*>(a.k + o) + c);
                pv[c] = __ldg(reinterpret_cast<const uint4*>(a.v + o) + c);
            }
        }
    };
    if (t_begin < t_end) prefetch(t_begin);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    constexpr int O_COL = 128;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t id_s = umma_idesc_bf16(TILE, TILE, false, false), id_o = umma_idesc_bf16(TILE, HD, false, true);
    uint32_t phase = 0;
    int h = -1;
    float scale = 0.0f;

    for (int t = t_begin; t < t_end; ++t) {
        const int th = t / npairs, pair = t - th * npairs;
        if (th != h) {
            // every thread is past the previous item's softmax (its second barrier), so the old bias is dead; the staging barrier
            // below publishes the new one
            h = th;
            scale = __expf(fminf(a.logit_scale[h], LOGIT_MAX));
            const float* bh = a.bias + static_cast<long long>(h) * N * N;
            for (int idx = tid; idx < N * N; idx += 128) sB[(idx / N) * (N + 1) + idx % N] = __ldg(bh + idx);
        }
        const int tok = ptok, region = pregion;
        const int bw = pair * 2 + g;
        sreg[tid] = region;
        {
            float q[HD], k[HD];
            if (tok >= 0) {
#pragma unroll
                for (int c = 0; c < 4; ++c) { unpack8(pq[c], q + 8 * c); unpack8(pk[c], k + 8 * c); }
                float nq, nk;
                const float iq = inv_norm32(q, nq), ik = inv_norm32(k, nk);
#pragma unroll
                for (int c = 0; c < HD; ++c) { q[c] *= iq; k[c] *= ik; }
            } else {
#pragma unroll
                for (int c = 0; c < HD; ++c) { q[c] = 0.0f; k[c] = 0.0f; }
            }
            stage_row32_hilo(sQ, tid, q);
            stage_row32_hilo(sK, tid, k);
#pragma unroll
            for (int c = 0; c < 4; ++c) st_tile8_raw(sV, tid, c, tok >= 0 ? pv[c] : make_uint4(0, 0, 0, 0));
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            issue_cosine_logits(tmem, smem_u32(sQ), smem_u32(sK), id_s);
            umma_commit(&bars[0]);
        }
        if (t + 1 < t_end) prefetch(t + 1);
        mbar_wait(&bars[0], phase);
        tc_fence_after();

        const float* brow = sB + (tok >= 0 ? n : 0) * (N + 1);
        float sv[SLOT];
        float mx = -INFINITY;
#pragma unroll
        for (int c0 = 0; c0 < SLOT; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(trow + g * SLOT + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                const int j = c0 + t;
                float sc = -INFINITY;
                if (tok >= 0 && j < N) {
                    sc = __uint_as_float(r[t]) * scale + brow[j];
                    if (sreg[g * SLOT + j] != region) sc += -200.0f;
                    mx = fmaxf(mx, sc);
                }
                sv[j] = sc;
            }
        }
        float sum = 0.0f;
#pragma unroll
        for (int j = 0; j < SLOT; ++j) {
            const float e = (tok >= 0 && j < N) ? __expf(sv[j] - mx) : 0.0f;
            sv[j] = e;
            sum += e;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) st_tile8(sP + g * TB, tid, c, sv + 8 * c);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
#pragma unroll
            for (int ks = 0; ks < TILE / 16; ++ks)
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
        // the next iteration's staging overwrites sQ / sK / sV / sreg: every thread is past the softmax (second barrier above) and
        // both MMAs have completed (bars[1]); the accumulators are re-issued only after the next iteration's first barrier
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward: 1-D grid of one CTA per SM, 512 threads; work items (head, window pair) are numbered head-major and CTA c owns the
// contiguous range [c T / G, (c + 1) T / G) (see the forward kernel); the bias / logit-scale gradients of a head are flushed
// when the CTA moves on to the next head.
// TMEM: S 0..127 | dP 128..255 | dV 256..287 | dK 288..351 | dQ 352..415
//
// One window pair is a serial chain (stage operands -> S, dP MMAs -> softmax backward -> dV, dK, dQ MMAs -> write out) and
// the operands of a pair fill most of the SM's shared memory, so a second resident CTA is not an option.  The kernel is
// latency bound, not throughput bound; it therefore (1) spreads every SIMT phase over 16 warps -- staging: 4 threads per
// token row, 16 bytes of q / k / v / dO each; softmax backward: 4 threads per row, 16 keys each; write-out: dq / dk / dv rows
// by different warps -- and (2) prefetches the rows of the NEXT pair into registers while the tensor core and the softmax
// phase work on the current one.
// ------------------------------------------------------------------------------------------------------------------
constexpr int BWD_THREADS = 512;

__global__ void __launch_bounds__(BWD_THREADS, 1) swin_attn_bwd_tc_kernel(SwinTcArgs a, int npairs) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + smem_align_pad(smem_raw);      // 1024-byte aligned, still a __shared__ pointer (LDS / STS, 32-bit addressing)
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TB;
    uint8_t* sV = sK + TB;
    uint8_t* sdO = sV + TB;
    uint8_t* sP = sdO + TB;
    uint8_t* sdS = sP + 2 * TB;
    uint8_t* sdSl = sdS + 2 * TB;                                  // low-order bf16 part of dS
    float* dbias_s = reinterpret_cast<float*>(sdSl + 2 * TB);      // [N][N + 1] (padded: conflict-free), used by the final flush only
    const int N = a.N;
    int* sreg = reinterpret_cast<int*>(dbias_s + N * (N + 1));     // [128] shift-mask region of the row's token, -1 = padding row
    float* sqn = reinterpret_cast<float*>(sreg + TILE);            // [128] |q|
    float* skn = sqn + TILE;                                       // [128] |k|
    float* sD = skn + TILE;                                        // [4][128] partial D_i
    float* sB = sD + 4 * TILE;                                     // [N][N + 1] position bias of this head (see the forward kernel)
    float* red = sB + N * (N + 1);                                 // [16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 16);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
    constexpr int TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 288, TM_DQ = 352;      // dK / dQ: 64 columns ([hi | lo] of B)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = npairs * a.heads;
    const int t_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * T / gridDim.x);
    const int t_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * T / gridDim.x);
    // staging view: 4 consecutive threads share token row srow, each moves the 16-byte chunk `part` (8 channels) of q, k, v, dO
    const int srow = tid >> 2, part = tid & 3, sg = srow >> 6, sn = srow & 63;
    // TMEM view: TMEM lane = row; the four warps with the same (warp & 3) split the row's 64 keys into quarters
    const int r = (warp & 3) * 32 + lane, g = r >> 6, n = r & 63, cq = warp >> 2;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    // the bias gradient of element (n, j) is owned by the thread with TMEM row n (either window slot) and key quarter j / 16 for
    // EVERY window pair this CTA visits, so it accumulates in registers (a shared-memory row per thread would put the 32 lanes
    // of a warp on one bank: rows are 64 floats apart) and is combined once at the end
    float dbacc[16];
    // P / dS / dSl hold, per row, the keys of the row's OWN window in key block g; the other block stays zero for the whole
    // kernel (it is the cross-window half of the 128 x 128 product), so it is cleared once, not once per pair
    for (int idx = tid; idx < 6 * TB / 16; idx += BWD_THREADS) reinterpret_cast<uint4*>(sP)[idx] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    int h = -1;
    float raw_ls = 0.0f, scale = 0.0f;
    const uint32_t id_s = umma_idesc_bf16(TILE, TILE, false, false);   // S / dP
    const uint32_t id_t = umma_idesc_bf16(TILE, HD, true, true);       // dV / dK : A = P^T (MN-major), B MN-major
    const uint32_t id_q = umma_idesc_bf16(TILE, HD, false, true);      // dQ      : A = dS (K-major), B = Kh MN-major
    const uint32_t id_t64 = umma_idesc_bf16(TILE, 2 * HD, true, true);
    const uint32_t id_q64 = umma_idesc_bf16(TILE, 2 * HD, false, true);
    float dscale_acc = 0.0f;
    uint32_t phase = 0;

    // prefetch registers (rows of the pair about to be staged)
    uint4 pq = make_uint4(0, 0, 0, 0), pk = pq, pv = pq, pdo = pq;
    int ptok = -1, pregion = 0;
    auto prefetch = [&](int t) {
        const int ph = t / npairs, pair = t - ph * npairs;
        ptok = window_token(a, pair * 2 + sg, sn, pregion);
        if (ptok >= 0) {
            const long long o = static_cast<long long>(ptok) * a.ld + ph * HD;
            const long long oc = static_cast<long long>(ptok) * a.ldc + ph * HD;
            pq = __ldg(reinterpret_cast<const uint4*>(a.q + o) + part);
            pk = __ldg(reinterpret_cast<const uint4*>(a.k + o) + part);
            pv = __ldg(reinterpret_cast<const uint4*>(a.v + o) + part);
            pdo = __ldg(reinterpret_cast<const uint4*>(a.dctx + oc) + part);
        } else {
            pq = pk = pv = pdo = make_uint4(0, 0, 0, 0);
        }
    };
    if (t_begin < t_end) prefetch(t_begin);

    // flush of one head: bias gradient of every window this CTA visited for it (the two window slots are combined in shared
    // memory first) and d(logit_scale)
    auto flush_head = [&]() {
        if (n < N) {
#pragma unroll
            for (int t = 0; t < 16; ++t)
                if (cq * 16 + t < N && dbacc[t] != 0.0f) atomicAdd(&dbias_s[n * (N + 1) + cq * 16 + t], dbacc[t]);
        }
        __syncthreads();
        float* dbias = a.dbias + static_cast<long long>(h) * N * N;
        for (int idx = tid; idx < N * N; idx += BWD_THREADS) {
            const float v = dbias_s[(idx / N) * (N + 1) + idx % N];
            if (v != 0.0f) atomicAdd(&dbias[idx], v);
        }
        dscale_acc = warp_sum(dscale_acc);
        if (lane == 0) red[warp] = dscale_acc;
        __syncthreads();
        if (tid == 0) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < BWD_THREADS / 32; ++w) t += red[w];
            atomicAdd(&a.dlogit_scale[h], raw_ls <= LOGIT_MAX ? t * scale : 0.0f);
        }
        __syncthreads();          // dbias_s and red are re-initialised for the next head right after this
    };

    for (int t = t_begin; t < t_end; ++t) {
        const int th = t / npairs, pair = t - th * npairs;
        if (th != h) {
            if (h >= 0) flush_head();
            h = th;
            raw_ls = a.logit_scale[h];
            scale = __expf(fminf(raw_ls, LOGIT_MAX));
            dscale_acc = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; ++i) dbacc[i] = 0.0f;
            for (int idx = tid; idx < N * (N + 1); idx += BWD_THREADS) dbias_s[idx] = 0.0f;
            const float* bh = a.bias + static_cast<long long>(h) * N * N;
            for (int idx = tid; idx < N * N; idx += BWD_THREADS) sB[(idx / N) * (N + 1) + idx % N] = __ldg(bh + idx);
            // published by the staging barrier below (dbias_s is next touched by a flush, sB by the softmax backward)
        }
        // ---- stage the pair: L2-normalise q, k (fp32), split into bf16 hi | lo, copy v and dO ----
        {
            float q[8], k[8];
            unpack8(pq, q);
            unpack8(pk, k);
            float sq = 0.0f, sk = 0.0f;
#pragma unroll
            for (int c = 0; c < 8; ++c) { sq = fmaf(q[c], q[c], sq); sk = fmaf(k[c], k[c], sk); }
            sq += __shfl_xor_sync(0xffffffffu, sq, 1); sq += __shfl_xor_sync(0xffffffffu, sq, 2);
            sk += __shfl_xor_sync(0xffffffffu, sk, 1); sk += __shfl_xor_sync(0xffffffffu, sk, 2);
            const float qn = fmaxf(sqrtf(sq), NORM_EPS), kn = fmaxf(sqrtf(sk), NORM_EPS);
            const float iq = ptok >= 0 ? 1.0f / qn : 0.0f, ik = ptok >= 0 ? 1.0f / kn : 0.0f;
            float ql[8], kl[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                q[c] *= iq; k[c] *= ik;
                ql[c] = q[c] - __bfloat162float(__float2bfloat16_rn(q[c]));
                kl[c] = k[c] - __bfloat162float(__float2bfloat16_rn(k[c]));
            }
            st_tile8(sQ, srow, part, q);
            st_tile8(sQ, srow, 4 + part, ql);
            st_tile8(sK, srow, part, k);
            st_tile8(sK, srow, 4 + part, kl);
            st_tile8_raw(sV, srow, part, pv);
            st_tile8_raw(sdO, srow, part, pdo);
            if (part == 0) {
                sreg[srow] = ptok >= 0 ? pregion : -1;
                sqn[srow] = qn;
                skn[srow] = kn;
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
        // rows of the next pair travel while the tensor core and the softmax phase work on this one
        if (t + 1 < t_end) prefetch(t + 1);

        const int region = sreg[r];
        const bool valid = region >= 0;
        const int bw = pair * 2 + g;
        const float lse = valid ? a.lse[(static_cast<long long>(bw) * a.heads + h) * N + n] : 0.0f;
        const float* brow = sB + (valid ? n : 0) * (N + 1) + cq * 16;
        int kreg[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) kreg[t] = sreg[g * SLOT + cq * 16 + t];
        mbar_wait(&bars[0], phase);
        phase ^= 1;
        tc_fence_after();

        // pass 1: probabilities of this thread's 16 keys and the partial D_i = sum_j P_ij dP_ij.  D is taken from the SAME P
        // and dP that form dS = P (dP - D) -- not from dO . O with the bf16-rounded O -- so the cancellation in (dP - D) is exact
        // to fp32 rounding; with O rounded to 8 bits the q / k / bias gradients of peaked rows were rounding noise.
        uint32_t rs[16], rp[16];
        tmem_ld_32x16(trow + TM_S + g * SLOT + cq * 16, rs);
        tmem_ld_32x16(trow + TM_DP + g * SLOT + cq * 16, rp);
        tmem_ld_wait();
        float pv16[16];
        float Dp = 0.0f;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int j = cq * 16 + t;
            float p = 0.0f;
            if (valid && j < N) {
                float sc = __uint_as_float(rs[t]) * scale + brow[t];
                if (kreg[t] != region) sc += -200.0f;
                p = __expf(sc - lse);
                Dp = fmaf(p, __uint_as_float(rp[t]), Dp);
            }
            pv16[t] = p;
        }
        sD[cq * TILE + r] = Dp;
        __syncthreads();
        const float Di = (sD[r] + sD[TILE + r]) + (sD[2 * TILE + r] + sD[3 * TILE + r]);
        {
            float dsv[16], dlo[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int j = cq * 16 + t;
                float ds = 0.0f;
                if (valid && j < N) {
                    ds = pv16[t] * (__uint_as_float(rp[t]) - Di);
                    dbacc[t] += ds;
                    dscale_acc = fmaf(ds, __uint_as_float(rs[t]), dscale_acc);
                }
                dsv[t] = ds;
                dlo[t] = ds - __bfloat162float(__float2bfloat16_rn(ds));   // dS = hi + lo (both bf16): see the dQ / dK passes below
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                st_tile8(sP + g * TB, r, cq * 2 + c, pv16 + 8 * c);
                st_tile8(sdS + g * TB, r, cq * 2 + c, dsv + 8 * c);
                st_tile8(sdSl + g * TB, r, cq * 2 + c, dlo + 8 * c);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            const uint32_t pa = smem_u32(sP), dsa = smem_u32(sdS), dla = smem_u32(sdSl);
            const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK), doa = smem_u32(sdO);
            // dV = P^T dO.  dK = dS^T Qh and dQ = dS Kh at ~16-bit operand precision (their results go through the
            // cancelling normalisation backward): the B operand covers the [hi | lo] halves of Qh / Kh (N = 64, columns
            // c and 32 + c are summed in the epilogue) and a second pass adds dS_lo x hi.
#pragma unroll
            for (int ks = 0; ks < TILE / 16; ++ks) {
                umma_bf16(tmem + TM_DV, umma_smem_desc_sw128(pa + ks * 2048, TB, 1024), umma_smem_desc_sw128(doa + ks * 2048, 8192, 1024), id_t, ks != 0);
                umma_bf16(tmem + TM_DK, umma_smem_desc_sw128(dsa + ks * 2048, TB, 1024), umma_smem_desc_sw128(qa + ks * 2048, 8192, 1024), id_t64, ks != 0);
                umma_bf16(tmem + TM_DQ, umma_smem_desc_sw128(dsa + (ks >> 2) * TB + (ks & 3) * 32, 16, 1024),
                          umma_smem_desc_sw128(ka + ks * 2048, 8192, 1024), id_q64, ks != 0);
            }
#pragma unroll
            for (int ks = 0; ks < TILE / 16; ++ks) {
                umma_bf16(tmem + TM_DK, umma_smem_desc_sw128(dla + ks * 2048, TB, 1024), umma_smem_desc_sw128(qa + ks * 2048, 8192, 1024), id_t, 1u);
                umma_bf16(tmem + TM_DQ, umma_smem_desc_sw128(dla + (ks >> 2) * TB + (ks & 3) * 32, 16, 1024),
                          umma_smem_desc_sw128(ka + ks * 2048, 8192, 1024), id_q, 1u);
            }
            umma_commit(&bars[0]);
        }
        mbar_wait(&bars[0], phase);
        phase ^= 1;
        tc_fence_after();
        // ---- write out: warps 0-3 dq, warps 4-7 dk, warps 8-11 dv (token row = TMEM lane) ----
        if (cq < 3) {
            int rg;
            const int tok = window_token(a, bw, n, rg);
            if (cq == 2) {
                uint32_t rv[32];
                tmem_ld_32x32(trow + TM_DV, rv);
                tmem_ld_wait();
                if (tok >= 0) {
                    float o[HD];
#pragma unroll
                    for (int c = 0; c < HD; ++c) o[c] = __uint_as_float(rv[c]);
                    store_row32(a.dv, a.ld, tok, h, o);
                }
            } else {
                // d q = (d qh - qh (qh . d qh)) / |q| with d qh = scale * (dS Kh)   (normalisation backward); same for k
                uint32_t hi[32], lo[32];
                tmem_ld_32x32(trow + (cq == 0 ? TM_DQ : TM_DK), hi);
                tmem_ld_32x32(trow + (cq == 0 ? TM_DQ : TM_DK) + HD, lo);
                tmem_ld_wait();
                if (tok >= 0) {
                    const uint8_t* tile = cq == 0 ? sQ : sK;
                    float xh[HD], o[HD];
#pragma unroll
                    for (int c = 0; c < 4; ++c) unpack8(ld_tile8_raw(tile, r, c), xh + 8 * c);
                    float dot = 0.0f;
#pragma unroll
                    for (int c = 0; c < HD; ++c) {
                        o[c] = (__uint_as_float(hi[c]) + __uint_as_float(lo[c])) * scale;
                        dot = fmaf(o[c], xh[c], dot);
                    }
                    const float inv = 1.0f / (cq == 0 ? sqn[r] : skn[r]);
#pragma unroll
                    for (int c = 0; c < HD; ++c) o[c] = (o[c] - xh[c] * dot) * inv;
                    store_row32(cq == 0 ? a.dq : a.dk, a.ld, tok, h, o);
                }
            }
        }
        tc_fence_before();
        __syncthreads();          // tiles and TMEM accumulators are reused by the next window pair
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

}  // namespace

bool swin_attention_tc_supported(int dtype, int head_dim, int window, long long ld, long long ldc, const void* q, const void* k,
                                 const void* v, const void* ctx) {
    if (dtype != KLAB_BF16 || head_dim != HD || window * window > SLOT) return false;
    if ((ld | ldc) % 8) return false;
    return ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(ctx)) & 15) == 0;
}

static SwinTcArgs make_args(int B, int res, int heads, int window, int shift, const void* q, const void* k, const void* v, long long ld,
                            long long ldc, const float* logit_scale, const float* bias, float* lse) {
    SwinTcArgs a{};
    a.q = static_cast<const __nv_bfloat16*>(q); a.k = static_cast<const __nv_bfloat16*>(k); a.v = static_cast<const __nv_bfloat16*>(v);
    a.ld = ld; a.ldc = ldc; a.B = B; a.res = res; a.heads = heads; a.w = window; a.shift = shift; a.N = window * window;
    a.nW = (res / window) * (res / window);
    a.logit_scale = logit_scale; a.bias = bias; a.lse = lse;
    return a;
}

int swin_attention_fwd_tc(cudaStream_t st, int B, int res, int heads, int window, int shift, const void* q, const void* k, const void* v,
                          long long ld, void* ctx, long long ldc, const float* logit_scale, const float* bias, float* lse) {
    SwinTcArgs a = make_args(B, res, heads, window, shift, q, k, v, ld, ldc, logit_scale, bias, lse);
    a.out = static_cast<__nv_bfloat16*>(ctx);
    const int N = a.N;
    const size_t smem = 1024 + 5 * TB + sizeof(float) * N * (N + 1) + sizeof(int) * TILE + 64;
    static size_t set = 0;
    if (smem > set) {
        KLAB_CHECK_CUDA(cudaFuncSetAttribute(swin_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        set = smem;
    }
    const int npairs = (B * a.nW + 1) / 2;
    int grid = 2 * sm_count();                                      // two resident CTAs per SM, equal shares of the work items
    if (grid > npairs * heads) grid = npairs * heads;
    swin_attn_fwd_tc_kernel<<<grid, 128, smem, st>>>(a, npairs);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

int swin_attention_bwd_tc(cudaStream_t st, int B, int res, int heads, int window, int shift, const void* q, const void* k, const void* v,
                          long long ld, const void* ctx, const void* dctx, long long ldc, void* dq, void* dk, void* dv,
                          const float* logit_scale, const float* bias, const float* lse, float* dbias, float* dlogit_scale) {
    SwinTcArgs a = make_args(B, res, heads, window, shift, q, k, v, ld, ldc, logit_scale, bias, const_cast<float*>(lse));
    a.ctx = static_cast<const __nv_bfloat16*>(ctx); a.dctx = static_cast<const __nv_bfloat16*>(dctx);
    a.dq = static_cast<__nv_bfloat16*>(dq); a.dk = static_cast<__nv_bfloat16*>(dk); a.dv = static_cast<__nv_bfloat16*>(dv);
    a.dbias = dbias; a.dlogit_scale = dlogit_scale;
    const int N = a.N;
    const size_t smem = 1024 + 10 * TB + 2 * sizeof(float) * N * (N + 1) + sizeof(int) * TILE + sizeof(float) * (6 * TILE + 16) + 64;
    static size_t set = 0;
    if (smem > set) {
        KLAB_CHECK_CUDA(cudaFuncSetAttribute(swin_attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        set = smem;
    }
    KLAB_CHECK_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * heads * N * N, st));
    KLAB_CHECK_CUDA(cudaMemsetAsync(dlogit_scale, 0, sizeof(float) * heads, st));
    const int npairs = (B * a.nW + 1) / 2;
    int grid = sm_count();                                           // one resident CTA per SM, equal shares of the work items
    if (grid > npairs * heads) grid = npairs * heads;
    swin_attn_bwd_tc_kernel<<<grid, BWD_THREADS, smem, st>>>(a, npairs);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

}  // namespace klab
