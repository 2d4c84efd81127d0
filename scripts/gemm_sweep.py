"""Development aid: time the tcgen05 GEMM over N-tile widths / K splits for a list of shapes (CUDA-graph replays, CUDA events),
to calibrate the cost model in csrc/gemm_tc.cu:pick_config.   python scripts/gemm_sweep.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from klab_multimodalmodel_b200 import _lib as L
from klab_multimodalmodel_b200 import ops as O

dev = torch.device("cuda", 0)


def timeit(M, N, K, a_mn, b_mn, od, **kw):
    per = (M * K + N * K) * 2 + M * N * (2 if od == torch.bfloat16 else 4)
    copies = max(1, min(6, (300 << 20) // per))
    As = [torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16() for _ in range(copies)]
    Bs = [torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16() for _ in range(copies)]
    Ds = [torch.zeros(M, N, device=dev, dtype=od) for _ in range(copies)]
    extra = {}
    if kw.get("res"):
        extra["residual"] = torch.randn(M, N, device=dev).bfloat16()
    if kw.get("aux_in"):
        extra["aux_in"] = torch.randn(M, N, device=dev).bfloat16()
    if kw.get("bias"):
        extra["bias"] = torch.randn(N, device=dev)
    if kw.get("aux_out"):
        extra["aux_out"] = torch.empty(M, N, device=dev, dtype=od)
    extra["act"] = kw.get("act", 0)
    extra["dropout_p"] = kw.get("p", 0.0)
    for i in range(2):
        O.gemm(As[0], Bs[0], M, N, K, a_mn=a_mn, b_mn=b_mn, out=Ds[0], **extra)
    torch.cuda.synchronize()
    iters = 6
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            O.gemm(As[i % copies], Bs[i % copies], M, N, K, a_mn=a_mn, b_mn=b_mn, out=Ds[i % copies], **extra)
    g.replay()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


SHAPES = [
    # M, N, K, a_mn, b_mn, out dtype, epilogue
    (6144, 1024, 4096, 0, 0, torch.bfloat16, {}),
    (6144, 1024, 4096, 0, 1, torch.bfloat16, {}),
    (6144, 4096, 1024, 0, 0, torch.bfloat16, dict(act=1, p=0.1)),
    (6144, 4096, 1024, 0, 1, torch.bfloat16, dict(act=3, p=0.1, aux_in=1)),
    (6144, 1024, 1024, 0, 0, torch.bfloat16, dict(res=1, p=0.1)),
    (6144, 3072, 1024, 0, 0, torch.bfloat16, {}),
    (2048, 1024, 1024, 0, 0, torch.bfloat16, {}),
    (2048, 3072, 1024, 0, 0, torch.bfloat16, {}),
    (2048, 1024, 4096, 0, 0, torch.bfloat16, dict(res=1, p=0.1)),
    (16384, 2048, 512, 0, 0, torch.bfloat16, dict(act=2, bias=1, aux_out=1)),
    (16384, 2048, 512, 0, 1, torch.bfloat16, dict(act=4, aux_in=1)),
    (16384, 512, 2048, 0, 0, torch.bfloat16, dict(bias=1)),
    (262144, 512, 128, 0, 0, torch.bfloat16, dict(act=2, bias=1, aux_out=1)),
    (1024, 1024, 6144, 1, 1, torch.float32, {}),
    (4096, 1024, 6144, 1, 1, torch.float32, {}),
    (1024, 1024, 2048, 1, 1, torch.float32, {}),
    (512, 2048, 16384, 1, 1, torch.float32, {}),
    (512, 512, 16384, 1, 1, torch.float32, {}),
    (128, 512, 262144, 1, 1, torch.float32, {}),
]

for M, N, K, a_mn, b_mn, od, kw in SHAPES:
    os.environ.pop("KLAB_GEMM_FORCE_BN", None)
    os.environ.pop("KLAB_GEMM_FORCE_SPLITS", None)
    auto = timeit(M, N, K, a_mn, b_mn, od, **kw)
    fl = 2.0 * M * N * K
    line = f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} {kw}: auto {auto:.1f}us ({fl / auto / 1e6:.0f} TF) |"
    splittable = od == torch.float32
    for bn in ((64, 128, 192, 256) if b_mn else (64, 96, 128, 160, 176, 192, 224, 256)):
        os.environ["KLAB_GEMM_FORCE_BN"] = str(bn)
        for sp in ((1, 2, 4, 8, 16) if splittable else (1,)):
            if sp > 1:
                os.environ["KLAB_GEMM_FORCE_SPLITS"] = str(sp)
            else:
                os.environ.pop("KLAB_GEMM_FORCE_SPLITS", None)
            t = timeit(M, N, K, a_mn, b_mn, od, **kw)
            line += f" bn{bn}" + (f"s{sp}" if splittable else "") + f" {t:.1f}"
    print(line, flush=True)
