"""GPU probe for the tcgen05 GEMM: correctness per operand-major / tile variant + first throughput numbers.

    python scripts/gpu_gemm_probe.py <variant>|all|perf

Each variant runs in its own process when driven by scripts/gpu_gemm_probe.sh, so a trap in one variant
does not poison the CUDA context of the others.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from klab_multimodalmodel_b200 import _lib as L


def run_gemm(a, b, a_mn, b_mn, M, N, K, out_dtype=torch.bfloat16, simt=False, **epi_kw):
    lib = L.lib()
    d = torch.empty(M, N, device="cuda", dtype=out_dtype)
    e = L.GemmEpilogue()
    e.alpha = 1.0
    e.out_dtype = L.BF16 if out_dtype == torch.bfloat16 else L.F32
    for k, v in epi_kw.items():
        setattr(e, k, v)
    fn = lib.klab_gemm_simt if simt else lib.klab_gemm
    in_dt = L.BF16 if a.dtype == torch.bfloat16 else L.F32
    L.check(fn(torch.cuda.current_stream().cuda_stream, in_dt, M, N, K, a.data_ptr(), a.stride(0), a_mn,
               b.data_ptr(), b.stride(0), b_mn, d.data_ptr(), d.stride(0), C.byref(e)))
    return d


def variant(a_mn, b_mn, M, N, K, tag):
    torch.manual_seed(0)
    A = torch.randn(M, K, device="cuda").bfloat16()          # logical A(m,k)
    B = torch.randn(N, K, device="cuda").bfloat16()          # logical B(n,k)
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    ref = A.float() @ B.float().t()
    d = run_gemm(a, b, a_mn, b_mn, M, N, K, out_dtype=torch.float32)
    torch.cuda.synchronize()
    err = (d - ref).abs().max().item()
    scale = ref.abs().max().item()
    ok = err <= 2e-3 * scale + 1e-3
    print(f"[{tag}] a_mn={a_mn} b_mn={b_mn} M={M} N={N} K={K}: max_abs_err={err:.4e} (ref max {scale:.3e}) {'OK' if ok else 'FAIL'}",
          flush=True)
    if not ok:
        bad = ((d - ref).abs() > 2e-3 * scale + 1e-3)
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print(f"   bad rows: n={rows.numel()} first={rows[:8].tolist()}  bad cols: n={cols.numel()} first={cols[:8].tolist()}")
        print("   d[0,:8]  ", d[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
    return ok


VARIANTS = {
    # name: (a_mn, b_mn, M, N, K)
    "kk_256": (0, 0, 512, 1024, 256),
    "kk_128": (0, 0, 256, 384, 192),
    "kk_64": (0, 0, 128, 64, 64),
    "kk_ragged": (0, 0, 333, 200, 136),
    "kn_256": (0, 1, 512, 1024, 256),
    "kn_ragged": (0, 1, 200, 136, 333),
    "nn_256": (1, 1, 512, 1024, 256),
    "nn_ragged": (1, 1, 136, 200, 333),
    "nk_128": (1, 0, 256, 384, 192),
}


def perf():
    lib = L.lib()
    for (M, N, K) in [(8192, 8192, 8192), (6144, 4096, 1024), (6144, 1024, 4096), (2048, 1024, 1024), (262144, 512, 128)]:
        for (a_mn, b_mn) in [(0, 0), (0, 1), (1, 1)]:
            A = torch.randn(K if a_mn else M, M if a_mn else K, device="cuda").bfloat16()
            B = torch.randn(K if b_mn else N, N if b_mn else K, device="cuda").bfloat16()
            for _ in range(3):
                run_gemm(A, B, a_mn, b_mn, M, N, K)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            s.record()
            for _ in range(iters):
                run_gemm(A, B, a_mn, b_mn, M, N, K)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / iters
            tf = 2.0 * M * N * K / ms / 1e9
            # cuBLAS for context
            if not a_mn and not b_mn:
                for _ in range(3):
                    torch.matmul(A, B.t())
                s.record()
                for _ in range(iters):
                    torch.matmul(A, B.t())
                e.record()
                torch.cuda.synchronize()
                ms2 = s.elapsed_time(e) / iters
                extra = f" | cuBLAS {2.0 * M * N * K / ms2 / 1e9:.1f} TFLOP/s"
            else:
                extra = ""
            print(f"[perf] M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: {ms * 1e3:.1f} us  {tf:.1f} TFLOP/s{extra}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    L.check(L.lib().klab_check_device())
    if which == "perf":
        perf()
    elif which == "all":
        ok = all([variant(*v, tag=k) for k, v in VARIANTS.items()])
        sys.exit(0 if ok else 1)
    else:
        sys.exit(0 if variant(*VARIANTS[which], tag=which) else 1)
