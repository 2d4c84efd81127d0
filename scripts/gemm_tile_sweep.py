"""Tile sweep of the large T5 GEMM shapes (development aid): per-launch microseconds inside a CUDA graph for every legal N tile of
both kernel kinds, operands rotated beyond L2.   python scripts/gemm_tile_sweep.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from klab_multimodalmodel_b200 import _lib as L
from klab_multimodalmodel_b200 import ops as O

dev = torch.device("cuda", 0)
lib = L.lib()
torch.manual_seed(0)
SHAPES = [(6144, 1024, 4096, False, False), (6144, 1024, 4096, False, True), (6144, 4096, 1024, False, False), (6144, 4096, 1024, False, True),
          (6144, 3072, 1024, False, False), (6144, 1024, 3072, False, True), (6144, 1024, 1024, False, False), (3072, 1024, 6144, True, True),
          (1024, 4096, 6144, True, True), (16384, 2048, 512, False, False), (16384, 512, 2048, False, False)]


def time_graph(fns, n=16):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fns[i % len(fns)]()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


for (M, N, K, a_mn, b_mn) in SHAPES:
    od = torch.float32 if (a_mn and b_mn) else torch.bfloat16
    per = (M * K + N * K) * 2 + M * N * (4 if od == torch.float32 else 2)
    copies = max(2, min(8, (300 << 20) // per))
    sets = [(torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16(), torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16(),
             torch.empty(M, N, device=dev, dtype=od)) for _ in range(copies)]
    fns = [(lambda A=A, B=B, D=D: O.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out=D)) for A, B, D in sets]
    row = []
    forces = [(-1, -1, -1)] + [(p, bn, 1) for p in (0, 1) for bn in (256, 224, 192, 176, 160, 144, 128, 96, 64)]
    for force in forces:
        lib.klab_gemm_set_force(*force)
        try:
            us = time_graph(fns)
            bn, sp, c2 = C.c_int(0), C.c_int(0), C.c_int(0)
            lib.klab_gemm_last_config(C.byref(bn), C.byref(sp), C.byref(c2))
            if force[1] < 0 or bn.value == force[1]:
                row.append((us, f"{'auto ' if force[1] < 0 else ''}p{c2.value}/bn{bn.value}/s{sp.value}"))
        finally:
            lib.klab_gemm_set_force(-1, -1, -1)
    auto = row[0]
    best = min(row)
    print(f"M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)}: auto {auto[1]} {auto[0]:.1f} us ({2.0 * M * N * K / auto[0] / 1e6:.0f} TF) | best {best[1]} {best[0]:.1f} us "
          f"({2.0 * M * N * K / best[0] / 1e6:.0f} TF) | " + "  ".join(f"{n}:{u:.1f}" for u, n in sorted(row[1:])[:6]), flush=True)
