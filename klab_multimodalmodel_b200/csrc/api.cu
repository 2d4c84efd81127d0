// C-ABI entry points (include/klab_b200.h) and host-side plumbing shared by all kernels.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "gemm.cuh"

namespace klab {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<int> g_sm_reserve{0};

// SMs the persistent kernels may fill: the device's SM count minus the reserve set by klab_set_sm_reserve (SMs left to a
// concurrently running collective, whose CTAs cannot be co-resident with a full-shared-memory CTA of ours).
int sm_count_physical() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

int sm_count() {
    const int n = sm_count_physical();
    const int r = g_sm_reserve.load(std::memory_order_relaxed);
    return n - r > 16 ? n - r : (n < 16 ? n : 16);
}

// Work-item counters of the dynamically scheduled persistent kernels (GEMM, attention): {next item, finished CTAs} per slot,
// zero between launches (the last CTA of a launch re-arms its slot, so CUDA graphs can replay it).  Launches take slots
// round-robin; two launches share a slot only when SCHED_SLOTS launches apart, long after the first has finished in any
// stream order this library produces.  Returns nullptr (= static striding) when the pool does not exist yet and cannot be
// created because the stream is capturing, or when dynamic distribution is off -- the default on a single GPU, where the static
// stride is ~1 us per launch cheaper; the data-parallel reducer switches it on (klab_set_dynamic_sched).
constexpr int SCHED_SLOTS = 8192;
static int* g_sched_base[16] = {};
static std::atomic<unsigned> g_sched_seq{0};

static std::atomic<int> g_dynamic_sched{-1};          // -1: KLAB_DYNAMIC_SCHED (default 0); 0 / 1: klab_set_dynamic_sched

int sched_slot_enabled() {
    int on = g_dynamic_sched.load(std::memory_order_relaxed);
    if (on < 0) {
        const char* e = getenv("KLAB_DYNAMIC_SCHED");
        on = e && e[0] == '1';
    }
    return on;
}

int* sched_slot(cudaStream_t stream) {
    if (!sched_slot_enabled()) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    if (!g_sched_base[dev]) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return nullptr;
        int* p = nullptr;
        if (cudaMalloc(&p, SCHED_SLOTS * 2 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (cudaMemset(p, 0, SCHED_SLOTS * 2 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); cudaFree(p); return nullptr; }
        g_sched_base[dev] = p;
    }
    return g_sched_base[dev] + 2 * (g_sched_seq.fetch_add(1, std::memory_order_relaxed) % SCHED_SLOTS);
}

// cuTensorMapEncodeTiled is fetched through the runtime so that the library carries no link-time
// dependency on libcuda.so (it must load, and export its symbols, on a box without a driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                      uint32_t box_cols) {
    EncodeTiledFn fn = encode_tiled_fn();
    // cuTensorMapEncodeTiled is a DRIVER call: it needs a current context in the calling thread.  The first klab call of a
    // fresh thread (autograd's backward thread: LMHeadLossFn.backward starts with a TMA-fed GEMM) may arrive before any runtime
    // call has bound the primary context there (CUDA_ERROR_INVALID_CONTEXT, 201) -- bind it once per thread.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        cudaFree(nullptr);
        ctx_bound = true;
    }
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
        return KLAB_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld * 2};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu box=%ux%u base=%p", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols, base);
        return KLAB_ERR_CUDA;
    }
    return KLAB_OK;
}

static int check_device_impl() {
    static int cached = -1;
    if (cached >= 0) return cached;
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        set_error("no CUDA device available: this library has no CPU fallback");
        return KLAB_ERR_UNSUPPORTED;
    }
    if (major != 10) {
        set_error("device compute capability %d.x is not supported: the kernels are built for sm_100a only", major);
        return KLAB_ERR_UNSUPPORTED;
    }
    cached = KLAB_OK;
    return cached;
}

}  // namespace klab

using namespace klab;

extern "C" {

int klab_abi_version(void) { return KLAB_ABI_VERSION; }
const char* klab_last_error(void) { return g_err; }
int klab_check_device(void) { return check_device_impl(); }
long long klab_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int klab_set_sm_reserve(int n_sms) {
    g_sm_reserve.store(n_sms < 0 ? 0 : n_sms, std::memory_order_relaxed);
    return KLAB_OK;
}
int klab_sm_budget(void) { return sm_count(); }
int klab_set_dynamic_sched(int on) {
    g_dynamic_sched.store(on ? 1 : 0, std::memory_order_relaxed);
    return KLAB_OK;
}

static klab_gemm_epilogue default_epilogue(const klab_gemm_epilogue* epi, int in_dtype) {
    klab_gemm_epilogue e;
    if (epi) {
        e = *epi;
    } else {
        memset(&e, 0, sizeof(e));
        e.alpha = 1.0f;
        e.out_dtype = in_dtype;
    }
    return e;
}

int klab_gemm(void* stream, int in_dtype, int M, int N, int K, const void* A, long long lda, int a_mn_major, const void* B,
              long long ldb, int b_mn_major, void* D, long long ldd, const klab_gemm_epilogue* epi) {
    if (int rc = check_device_impl()) return rc;
    const klab_gemm_epilogue e = default_epilogue(epi, in_dtype);
    if (in_dtype == KLAB_BF16)
        return gemm_tc_launch(static_cast<cudaStream_t>(stream), M, N, K, A, lda, a_mn_major, B, ldb, b_mn_major, D, ldd, e);
    return gemm_simt_launch(static_cast<cudaStream_t>(stream), in_dtype, M, N, K, A, lda, a_mn_major, B, ldb, b_mn_major, D, ldd, e);
}

int klab_gemm_set_force(int cta2, int bn, int splits) {
    gemm_set_force(cta2, bn, splits);
    return KLAB_OK;
}
int klab_gemm_last_config(int* bn, int* splits, int* cta2) {
    gemm_last_config(bn, splits, cta2);
    return KLAB_OK;
}

int klab_gemm_simt(void* stream, int in_dtype, int M, int N, int K, const void* A, long long lda, int a_mn_major,
                   const void* B, long long ldb, int b_mn_major, void* D, long long ldd, const klab_gemm_epilogue* epi) {
    if (int rc = check_device_impl()) return rc;
    const klab_gemm_epilogue e = default_epilogue(epi, in_dtype);
    return gemm_simt_launch(static_cast<cudaStream_t>(stream), in_dtype, M, N, K, A, lda, a_mn_major, B, ldb, b_mn_major, D, ldd, e);
}

}  // extern "C"
