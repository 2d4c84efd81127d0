"""Fixed cost of one tcgen05 GEMM launch inside a CUDA graph (development aid): back-to-back launches of small shapes, per-launch
microseconds, for the library's own choice and for pinned N tiles.   python scripts/gemm_overhead_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from klab_multimodalmodel_b200 import _lib as L
from klab_multimodalmodel_b200 import ops as O

dev = torch.device("cuda", 0)
lib = L.lib()
torch.manual_seed(0)


def time_graph(fn, n=200):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


SHAPES = [(128, 64, 64, False, False), (256, 768, 768, False, False), (256, 2304, 768, False, False), (256, 3072, 768, False, False),
          (256, 768, 3072, False, False), (256, 32128, 768, False, False), (2048, 1024, 1024, False, True), (2048, 1024, 1024, False, False),
          (2048, 3072, 1024, False, False), (1024, 1024, 2048, True, True)]
for (M, N, K, a_mn, b_mn) in SHAPES:
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    od = torch.float32 if (a_mn and b_mn) else torch.bfloat16
    D = torch.empty(M, N, device=dev, dtype=od)
    row = []
    for force in [(-1, -1, -1), (0, 16, -1), (0, 32, -1), (0, 64, -1), (0, 128, -1), (0, 256, -1), (1, 64, -1), (1, 128, -1), (1, 256, -1), (0, -1, 1), (1, -1, 1)]:
        lib.klab_gemm_set_force(*force)
        try:
            us = time_graph(lambda: O.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out=D))
            import ctypes as C
            bn, sp, c2 = C.c_int(0), C.c_int(0), C.c_int(0)
            lib.klab_gemm_last_config(C.byref(bn), C.byref(sp), C.byref(c2))
            row.append(f"{force}->bn{bn.value}/s{sp.value}/p{c2.value}:{us:.1f}")
        except RuntimeError as ex:
            row.append(f"{force}:ERR")
        finally:
            lib.klab_gemm_set_force(-1, -1, -1)
    print(f"M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} PDL={os.environ.get('KLAB_PDL', '1')}: " + "  ".join(row), flush=True)
# an empty-ish kernel for reference: the norm kernel on one row
x = torch.randn(256, 768, device=dev).bfloat16()
g = torch.ones(768, device=dev)
print("rmsnorm 256x768 per launch: %.1f us" % time_graph(lambda: O.rmsnorm_fwd(x, g, 1e-6, save_stats=False)))
y = torch.empty_like(x)
print("cast 256x768 per launch: %.1f us" % time_graph(lambda: O.cast(x, torch.bfloat16, out=y)))
