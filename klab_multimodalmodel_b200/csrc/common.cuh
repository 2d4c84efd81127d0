// Common device/host helpers for the sm_100a kernels: PTX wrappers (mbarrier, TMA, tcgen05/TMEM),
// UMMA descriptors, error plumbing.  Everything here is written for B200 only.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

#include "../../include/klab_b200.h"

namespace klab {

// ------------------------------------------------------------------------------------------------
// Error plumbing (C-ABI functions never throw; they return a status and keep the last message).
// ------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define KLAB_CHECK_CUDA(expr)                                                                \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::klab::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return KLAB_ERR_CUDA;                                                    \
        }                                                                                    \
    } while (0)

#define KLAB_REQUIRE(cond, ...)                                                              \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            ::klab::set_error(__VA_ARGS__);                                                  \
            return KLAB_ERR_INVALID;                                                 \
        }                                                                                    \
    } while (0)

#define KLAB_LAUNCH_CHECK() KLAB_CHECK_CUDA(cudaGetLastError())

inline size_t dtype_size(int dt) { return dt == KLAB_BF16 ? 2 : 4; }

int sm_count();             // SMs the persistent kernels may fill (physical count minus klab_set_sm_reserve)
int sm_count_physical();
int sched_slot_enabled();                  // 1 while dynamic work distribution is switched on
int* sched_slot(cudaStream_t stream);      // {next item, finished CTAs} counters of one dynamically scheduled launch, or nullptr

// Dynamic work distribution of a persistent kernel (device side).  A CTA's FIRST item is static (its block index: no atomic
// round trip in front of the first load of every launch -- that alone cost ~2 us x 1 200 GEMMs per step); later items come
// from the counter, starting at gridDim.x.  `claim` is called by ONE thread per CTA; without a counter it is the static
// stride blockIdx.x + i * gridDim.x.  If a concurrent collective holds some SMs, the CTAs that start late process only their
// one static item; everything else has been taken by the CTAs that were resident.
struct WorkClaim {
    int* ctr;
    int next_static;
    __device__ __forceinline__ void init(int* c) { ctr = c; next_static = blockIdx.x + gridDim.x; ticket = -1; }
    __device__ __forceinline__ int first() const { return blockIdx.x; }
    __device__ __forceinline__ int claim() {
        if (ctr) return static_cast<int>(gridDim.x) + atomicAdd(&ctr[0], 1);
        const int v = next_static;
        next_static += gridDim.x;
        return v;
    }
    // After the CTA's last claim (the claiming thread has SEEN a result past the end, so all its claims have been performed):
    // count this CTA as finished; the last one re-arms the counters.  Split in two so that the atomic's round trip can overlap
    // the CTA's remaining work: finish_begin() right after the last claim, finish_end() at the very end.  No fence is needed:
    // the next user of the slot is a later launch.
    int ticket;
    __device__ __forceinline__ void finish_begin() { ticket = ctr ? atomicAdd(&ctr[1], 1) : -1; }
    __device__ __forceinline__ void finish_end() {
        if (ctr && ticket == static_cast<int>(gridDim.x) - 1) {
            ctr[0] = 0;
            ctr[1] = 0;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Scalar conversion helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float load_as_f32(const void* p, size_t idx, int dt) {
    return dt == KLAB_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx])
                           : reinterpret_cast<const float*>(p)[idx];
}
__device__ __forceinline__ void store_from_f32(void* p, size_t idx, int dt, float v) {
    if (dt == KLAB_BF16) reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(p)[idx] = v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exact (erf) GELU, as HF `hidden_act="gelu"` (configuration_swinv2.py:69).  erfc is evaluated with the Abramowitz-Stegun
// 7.1.26 rational form (|error| <= 1.5e-7, i.e. fp32 rounding level): erfc(z) = poly(t) exp(-z^2), t = 1 / (1 + p z), z >= 0.
// It is branch free and costs one MUFU.RCP + one MUFU.EX2, which matters because the GEMM epilogue that applies it is
// instruction-issue bound (libdevice erff is ~40 instructions with two divergent branches).  GELU' reuses the same
// exponential: phi(x) = exp(-x^2 / 2) / sqrt(2 pi) and exp(-z^2) with z = |x| / sqrt(2) are the same number.
__device__ __forceinline__ float rcp_approx(float x) {     // one MUFU.RCP (__frcp_rn is the ~11-instruction IEEE sequence)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {     // one MUFU.EX2
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void gelu_terms(float x, float& half_erfc, float& u) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));             // argument in [1, inf): relative error ~1e-7
    u = ex2_approx(z * z * -1.4426950408889634f);                      // exp(-z^2)
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    half_erfc = 0.5f * poly * t * u;                       // 0.5 * erfc(|x| / sqrt 2) = 1 - Phi(|x|)
}
// gelu(x) and gelu'(x) from ONE erfc / exp evaluation (KLAB_ACT_GELU_SAVE_GRAD)
__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& gp) {
    float q, u;
    gelu_terms(x, q, u);
    const float cdf = x >= 0.0f ? 1.0f - q : q;
    g = x * cdf;
    gp = fmaf(x * 0.39894228040143267794f, u, cdf);
}
#ifdef KLAB_EXACT_GELU
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}
#else
__device__ __forceinline__ float gelu_erf(float x) {
    float q, u;
    gelu_terms(x, q, u);
    return x * (x >= 0.0f ? 1.0f - q : q);                 // x * Phi(x)
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
    float q, u;
    gelu_terms(x, q, u);
    const float cdf = x >= 0.0f ? 1.0f - q : q;
    return fmaf(x * 0.39894228040143267794f, u, cdf);      // Phi(x) + x phi(x)
}
#endif

// Counter-based dropout RNG: keep(element) is a pure function of (seed, element index), so the backward pass regenerates
// the forward mask without storing it.  The generator is cheap on purpose (the kernels that apply it are issue bound):
// one 64-bit mix of the seed per THREAD gives two 32-bit keys; one 32-bit hash per PAIR of elements gives two 16-bit
// uniforms.  keep probability = thr16 / 65536 and the survivors are scaled by exactly 65536 / thr16, so E[mask] = 1.
struct DropKey {
    uint32_t k0, k1;
    uint32_t thr16;       // keep iff 16-bit uniform < thr16
    float inv_keep;
    bool on;
};
__host__ __device__ inline uint32_t dropout_thr16(float p) {
    const double keep = 1.0 - static_cast<double>(p);
    uint32_t t = static_cast<uint32_t>(keep * 65536.0 + 0.5);
    return t > 65536u ? 65536u : (t < 1u ? 1u : t);
}
__device__ __forceinline__ DropKey make_drop_key(uint64_t seed, float p) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    DropKey k;
    k.k0 = static_cast<uint32_t>(z);
    k.k1 = static_cast<uint32_t>(z >> 32);
    k.on = p > 0.0f;
    k.thr16 = k.on ? dropout_thr16(p) : 65536u;
    k.inv_keep = k.on ? 65536.0f / static_cast<float>(k.thr16) : 1.0f;
    return k;
}
// 32 random bits for the element pair (2*pair, 2*pair + 1)
__device__ __forceinline__ uint32_t drop_hash_pair(const DropKey& k, uint64_t pair) {
    uint32_t x = static_cast<uint32_t>(pair) ^ k.k0;
    x += static_cast<uint32_t>(pair >> 32) * 0x9E3779B9u;
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16; x ^= k.k1;
    x *= 0x9E3779B1u;
    x ^= x >> 15;
    return x;
}
// multiplier of one element: 0 or 1 / keep
__device__ __forceinline__ float dropout_mult(const DropKey& k, uint64_t idx) {
    const uint32_t h = drop_hash_pair(k, idx >> 1);
    const uint32_t u = (idx & 1) ? (h >> 16) : (h & 0xFFFFu);
    return u < k.thr16 ? k.inv_keep : 0.0f;
}
// v[i] *= multiplier(base + i) for CH consecutive elements (CH even); one hash per pair when base is even
template <int CH>
__device__ __forceinline__ void dropout_apply_run(const DropKey& k, uint64_t base, float (&v)[CH]) {
    if ((base & 1) == 0) {
#pragma unroll
        for (int i = 0; i < CH; i += 2) {
            const uint32_t h = drop_hash_pair(k, (base + i) >> 1);
            v[i] *= (h & 0xFFFFu) < k.thr16 ? k.inv_keep : 0.0f;
            v[i + 1] *= (h >> 16) < k.thr16 ? k.inv_keep : 0.0f;
        }
    } else {
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] *= dropout_mult(k, base + i);
    }
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// Bytes to add to a dynamic-shared-memory base to reach the next 1024-byte boundary (SWIZZLE_128B tiles need it).  Adding the
// pad to the __shared__ ARRAY keeps the pointer in the shared address space: rounding it up through uintptr_t makes it a generic
// pointer, and every tile access becomes a generic LD / ST with 64-bit address arithmetic instead of LDS / STS.
__device__ __forceinline__ uint32_t smem_align_pad(const void* base) { return (1024u - (smem_u32(base) & 1023u)) & 1023u; }


__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            printf("klab: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// Same, for waits with slack (a producer waiting for a free ring slot, the MMA issuer waiting for a drained accumulator): back
// off between polls so that the polling does not take issue slots from the warps doing the work (the fused GEMM epilogues are
// issue bound; ncu attributed ~9 % of all issued instructions of a GELU GEMM to the spin loops of two waiting threads).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(40);
        if (++spins > (1u << 24)) {
            printf("klab: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// TMA 2D tile load: global (tensor map) -> shared, completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256); the address must be 32-byte aligned
__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
                 "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start while its
// predecessor is still running; it must not touch global memory the predecessor reads or writes before pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// --- TMEM allocation ---
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// --- CTA pair (cta_group::2): two SMs of one TPC run one UMMA of M = 256; each CTA stages its 128 rows of A and HALF of the B
// tile, which halves the L2 -> SM operand traffic per SM.  A shared-memory address with bit 24 cleared names the same offset
// in the EVEN (leader) CTA of the pair (CUTLASS: Sm100MmaPeerBitMask).
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load issued by either CTA of a pair; the transaction bytes are credited to the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
// arrive (count 1) on the mbarrier at this offset in the leader CTA of the pair (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
// --- cluster-scope shared-memory traffic between the two CTAs of a pair (dynamic work queue of the pair GEMM) ---
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t cta_rank) {     // same offset in CTA `cta_rank` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_shared_cluster_b32(uint32_t cluster_addr, int v) {
    asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
// arrive (count 1) on an mbarrier of another CTA of the cluster; release at cluster scope: what this thread wrote (or read) before is
// ordered before the arrival as seen by a cluster-scope acquire wait
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader_release(uint64_t* bar) {
    mbar_arrive_cluster(smem_u32(bar) & PEER_BIT_MASK);
}
// wait with cluster-scope acquire (the arrival may come from the peer CTA and publish data it stored into this CTA's shared memory)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred P;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (++spins > (1u << 26)) {
            printf("klab: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs of the pair once all previously issued pair-MMAs have completed
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane_base + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp in CUTLASS 4.x; restated, not included)
// ------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B, version 1 (Blackwell).
//   K-major  tile: rows of 128 B (64 bf16 of K); 8-row groups 1024 B apart -> SBO = 1024, LBO unused.
//   MN-major tile: rows of 128 B (64 bf16 of M/N) indexed by k; 8-k groups 1024 B apart -> SBO = 1024;
//                  the next 64 elements of M/N live `lbo_bytes` further on -> LBO.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);           // start address  [0,14)
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;       // leading byte offset [16,30)
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;       // stride byte offset  [32,46)
    d |= static_cast<uint64_t>(1) << 46;                                // descriptor version = 1
    d |= static_cast<uint64_t>(2) << 61;                                // layout type: SWIZZLE_128B
    return d;
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulators.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                                 // c_format = F32
           | (1u << 7)                               // a_format = BF16
           | (1u << 10)                              // b_format = BF16
           | (static_cast<uint32_t>(a_mn_major) << 15)
           | (static_cast<uint32_t>(b_mn_major) << 16)
           | (static_cast<uint32_t>(n >> 3) << 17)
           | (static_cast<uint32_t>(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// Host: TMA tensor maps (driver entry point fetched at run time: the library has no libcuda link
// dependency, so it loads -- and exports its symbols -- on a box without a GPU driver).
// ------------------------------------------------------------------------------------------------
// 2-D bf16 row-major tensor [rows, cols] with row stride `ld` elements; box = [box_rows, box_cols];
// 128-byte swizzle (box_cols * 2 bytes must be <= 128).
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);

}  // namespace klab
