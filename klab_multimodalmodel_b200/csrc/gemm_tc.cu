// K5: bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed smem ring).
//
//   D[M,N] = epilogue( alpha * A(m,k) B(n,k) ),  A/B bf16, fp32 accumulation in TMEM.
//
// Replaces the cuBLAS GEMM + separate bias / activation / residual kernels behind every nn.Linear on the
// path (HF/models/swinv2/modeling_swinv2.py:535,578-579,591,385; HF/models/t5/modeling_t5.py:93-102,
// 178-181,338,1110) for forward (both operands K-major), dgrad (B MN-major) and wgrad (A and B MN-major).
//
// Structure (persistent, one CTA per SM, 192 threads):
//   warp 0      : TMA producer   -- cp.async.bulk.tensor tiles into a STAGES-deep smem ring (SWIZZLE_128B)
//   warp 1      : MMA issuer     -- one elected lane issues tcgen05.mma (128 x BN x 16), commits to mbarriers
//   warps 2..5  : epilogue       -- tcgen05.ld the 128 x BN fp32 tile out of TMEM, fused epilogue, global stores
// Two TMEM accumulator stages (2 x BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#include "gemm.cuh"

namespace klab {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int A_TILE_BYTES = BM * BK * 2;   // 16 KiB

template <int BN> struct TileCfg {
    static constexpr int B_TILE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = 2 * BN;            // 128 / 256 / 512: powers of two
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    void* __restrict__ D, long long ldd, int M, int N, int K, int splits, klab_gemm_epilogue epi) {
    using Cfg = TileCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m = (M + BM - 1) / BM;
    const int num_n = (N + BN - 1) / BN;
    const int num_k = (K + BK - 1) / BK;
    // split-K (wgrad of tall-skinny activations): a work item is (output tile, K range); partial tiles are reduced with
    // fp32 red.global.add into a zero-initialised (or accumulating) D
    const int kps = (num_k + splits - 1) / splits;
    const int num_tiles = num_m * num_n * splits;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
                const int tile = item / splits, split = item - tile * splits;
                const int m0 = (tile / num_n) * BM;
                const int n0 = (tile % num_n) * BN;
                const int kb0 = split * kps, kb1 = min(num_k, kb0 + kps);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + A_TILE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    const int k0 = kb * BK;
                    if constexpr (!A_MN) {
                        tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);            // box {64 k, 128 m}
                    } else {
#pragma unroll
                        for (int i = 0; i < BM / 64; ++i)                                 // box {64 m, 64 k}
                            tma_load_2d(sa + i * 8192, &tmap_a, &full_bar[stage], m0 + i * 64, k0);
                    }
                    if constexpr (!B_MN) {
                        tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);            // box {64 k, BN n}
                    } else {
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i)                                 // box {64 n, 64 k}
                            tma_load_2d(sb + i * 8192, &tmap_b, &full_bar[stage], n0 + i * 64, k0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
                const int split = item % splits;
                const int kb0 = split * kps, kb1 = min(num_k, kb0 + kps);
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = A_MN ? umma_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                                 : umma_smem_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t db = B_MN ? umma_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                                 : umma_smem_desc_sw128(sb + k * 32, 16, 1024);
                        umma_bf16(d_tmem, da, db, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);           // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full_bar[acc]);             // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int sub = warp & 3;                              // TMEM sub-partition this warp may access
        const EpiDropout dr = make_dropout(epi.dropout_p);
        if (dr.on && epi.dropout_seed_ptr) epi.dropout_seed += *epi.dropout_seed_ptr;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
            const int tile = item / splits;
            const int m0 = (tile / num_n) * BM;
            const int n0 = (tile % num_n) * BN;
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const long long row = static_cast<long long>(m0) + sub * 32 + lane;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(t_row + c * 32, r);
                tmem_ld_wait();
                const int col0 = n0 + c * 32;
                const int nvalid = min(32, N - col0);
                if (row < M && nvalid > 0) {
                    if (splits > 1) {
                        float* dst = reinterpret_cast<float*>(D) + row * ldd + col0;
                        if (nvalid == 32) {                     // 16-byte aligned (ldd % 4 == 0, col0 % 32 == 0): vector reds
#pragma unroll
                            for (int i = 0; i < 32; i += 4)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i),
                                             "f"(__uint_as_float(r[i]) * epi.alpha), "f"(__uint_as_float(r[i + 1]) * epi.alpha),
                                             "f"(__uint_as_float(r[i + 2]) * epi.alpha), "f"(__uint_as_float(r[i + 3]) * epi.alpha)
                                             : "memory");
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (i < nvalid) atomicAdd(dst + i, __uint_as_float(r[i]) * epi.alpha);
                        }
                    } else {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                        epilogue_apply_store<32>(epi, dr, v, row, col0, nvalid, N, D, ldd);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int BN, bool A_MN, bool B_MN>
int launch_cfg(cudaStream_t stream, int M, int N, int K, int splits, const void* A, long long lda, const void* B, long long ldb,
               void* D, long long ldd, const klab_gemm_epilogue& epi) {
    using Cfg = TileCfg<BN>;
    CUtensorMap ta, tb;
    int rc;
    // A: K-major -> tensor [M rows, K cols], box [128, 64];  MN-major -> tensor [K rows, M cols], box [64, 64]
    rc = A_MN ? make_tmap_2d_bf16(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, 64)
              : make_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, 64);
    if (rc) return rc;
    rc = B_MN ? make_tmap_2d_bf16(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, 64)
              : make_tmap_2d_bf16(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN, 64);
    if (rc) return rc;
    auto kern = gemm_bf16_tc_kernel<BN, A_MN, B_MN>;
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
        KLAB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    if (splits > 1 && !epi.accumulate) {
        if (ldd == N) KLAB_CHECK_CUDA(cudaMemsetAsync(D, 0, sizeof(float) * static_cast<size_t>(M) * N, stream));
        else KLAB_CHECK_CUDA(cudaMemset2DAsync(D, sizeof(float) * ldd, 0, sizeof(float) * N, M, stream));
    }
    const int items = tiles * splits;
    const int grid = items < sm_count() ? items : sm_count();
    kern<<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, D, ldd, M, N, K, splits, epi);
    KLAB_LAUNCH_CHECK();
    count_launch();
    return KLAB_OK;
}

template <int BN>
int launch_major(cudaStream_t stream, int M, int N, int K, int splits, const void* A, long long lda, int a_mn, const void* B,
                 long long ldb, int b_mn, void* D, long long ldd, const klab_gemm_epilogue& epi) {
    if (!a_mn && !b_mn) return launch_cfg<BN, false, false>(stream, M, N, K, splits, A, lda, B, ldb, D, ldd, epi);
    if (!a_mn && b_mn) return launch_cfg<BN, false, true>(stream, M, N, K, splits, A, lda, B, ldb, D, ldd, epi);
    if (a_mn && !b_mn) return launch_cfg<BN, true, false>(stream, M, N, K, splits, A, lda, B, ldb, D, ldd, epi);
    return launch_cfg<BN, true, true>(stream, M, N, K, splits, A, lda, B, ldb, D, ldd, epi);
}

// Choose the N tile and the K split with a small time model (microseconds), calibrated on B200:
//   per k-block (64 deep) MMA time of a 128 x BN tile: BN/256 * 0.27 us (128 x 256 x 16 UMMA = 128 cycles), narrower tiles
//   lose a little to shared-memory bandwidth; ~2 us of prologue / epilogue per tile; split-K adds M*N*splits fp32 reds
//   (~0.2 elements/ns measured with red.global.add.v4.f32) and a memset.
void pick_config(int M, int N, int K, bool can_split, int* bn_out, int* splits_out) {
    const int sms = sm_count();
    const int num_m = (M + BM - 1) / BM, num_k = (K + BK - 1) / BK;
    const int cand[3] = {256, 128, 64};
    const double t_k[3] = {0.27, 0.145, 0.09};
    double best = 1e30;
    *bn_out = 64;
    *splits_out = 1;
    for (int i = 0; i < 3; ++i) {
        const int num_n = (N + cand[i] - 1) / cand[i];
        const long long tiles = 1ll * num_m * num_n;
        const double waste = double(num_n * cand[i]) / double(N);            // MMA work on padding columns is still paid
        for (int s = 1; s <= (can_split ? 128 : 1); s *= 2) {
            if (s > 1 && num_k / s < 8) break;
            const int kps = (num_k + s - 1) / s;
            const long long waves = (tiles * s + sms - 1) / sms;
            double t = waves * (kps * t_k[i] * (waste > 1.5 ? 1.0 : 1.0) + 2.0);
            if (s > 1) t += double(M) * N * s / 200.0e3 + 2.0;
            if (t < best - 1e-9) { best = t; *bn_out = cand[i]; *splits_out = s; }
        }
    }
    if (*splits_out > 1) {                                                   // no empty K ranges
        const int kps = (num_k + *splits_out - 1) / *splits_out;
        *splits_out = (num_k + kps - 1) / kps;
    }
}

}  // namespace

int gemm_tc_launch(cudaStream_t stream, int M, int N, int K, const void* A, long long lda, int a_mn, const void* B,
                   long long ldb, int b_mn, void* D, long long ldd, const klab_gemm_epilogue& epi) {
    KLAB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    KLAB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm(bf16): lda=%lld / ldb=%lld must be multiples of 8", lda, ldb);
    KLAB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                 "gemm(bf16): operand base pointers must be 16-byte aligned");
    // split-K needs a linear epilogue into an fp32 output (weight gradients)
    const bool can_split = epi.act == KLAB_ACT_NONE && !epi.bias && !epi.residual && !epi.aux_out && epi.dropout_p == 0.0f &&
                           epi.out_dtype == KLAB_F32 && ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0;
    int bn, splits;
    pick_config(M, N, K, can_split, &bn, &splits);
    switch (bn) {
        case 256: return launch_major<256>(stream, M, N, K, splits, A, lda, a_mn, B, ldb, b_mn, D, ldd, epi);
        case 128: return launch_major<128>(stream, M, N, K, splits, A, lda, a_mn, B, ldb, b_mn, D, ldd, epi);
        default: return launch_major<64>(stream, M, N, K, splits, A, lda, a_mn, B, ldb, b_mn, D, ldd, epi);
    }
}

}  // namespace klab
