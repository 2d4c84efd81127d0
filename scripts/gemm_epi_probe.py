"""Development aid: the epilogue-heavy GEMM signatures of workload 2a, each launched once inside a cudaProfilerStart/Stop range
(for `ncu --profile-from-start off --set full --import-source on`), plus CUDA-event timings.
python scripts/gemm_epi_probe.py [--time]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from klab_multimodalmodel_b200 import _lib as L
from klab_multimodalmodel_b200 import ops as O

ap = argparse.ArgumentParser()
ap.add_argument("--time", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
seedp = torch.zeros(1, dtype=torch.int64, device=dev)

CASES = {
    # name: (M, N, K, a_mn, b_mn, kwargs builder)
    "swin_fc1_dgrad_gelu_bwd": (16384, 2048, 512, False, True, lambda M, N: dict(act=L.ACT_GELU_BWD, aux_in=torch.randn(M, N, device=dev).bfloat16())),
    "swin_fc1_fwd_gelu": (16384, 2048, 512, False, False, lambda M, N: dict(act=L.ACT_GELU, bias=torch.randn(N, device=dev), aux_out=torch.empty(M, N, device=dev, dtype=torch.bfloat16))),
    "swin_fc1_plain": (16384, 2048, 512, False, False, lambda M, N: dict()),
    "t5_wi_dgrad_relu_bwd_drop": (6144, 4096, 1024, False, True, lambda M, N: dict(act=L.ACT_RELU_BWD, aux_in=torch.randn(M, N, device=dev).bfloat16(), dropout_p=0.1, seed=5, seed_ptr=seedp)),
    "t5_wi_fwd_relu_drop": (6144, 4096, 1024, False, False, lambda M, N: dict(act=L.ACT_RELU, dropout_p=0.1, seed=5, seed_ptr=seedp)),
    "t5_wi_plain": (6144, 4096, 1024, False, False, lambda M, N: dict()),
    "t5_wo_res_drop": (6144, 1024, 4096, False, False, lambda M, N: dict(residual=torch.randn(M, N, device=dev).bfloat16(), dropout_p=0.1, seed=5, seed_ptr=seedp)),
    "dec_small_plain": (2048, 1024, 1024, False, True, lambda M, N: dict()),
    "dec_wgrad_small": (1024, 1024, 2048, True, True, lambda M, N: dict(out_dtype=torch.float32)),
}
runs = []
for name, (M, N, K, a_mn, b_mn, mk) in CASES.items():
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    kw = mk(M, N)
    od = kw.pop("out_dtype", torch.bfloat16)
    D = torch.empty(M, N, device=dev, dtype=od)
    runs.append((name, (lambda A=A, B=B, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, kw=kw, D=D: O.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out=D, **kw)), 2.0 * M * N * K))

for _, fn, _ in runs:
    fn(); fn()
torch.cuda.synchronize()
if a.time:
    for name, fn, fl in runs:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            fn()
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) / 20 * 1e3
        print(f"{name:30s} {us:8.1f} us  {fl / us / 1e6:7.0f} TFLOP/s")
else:
    torch.cuda.profiler.start()
    for _, fn, _ in runs:
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
