"""The GEMM signatures that carry the most time in one step of bench workload 2a, once each inside a cudaProfilerStart/Stop
range, for `ncu --set full --profile-from-start off` (profiles/r02_gemm_top_*).  Operands exceed L2 between launches (every
signature has its own operand set and the sets are touched in rotation), epilogues as in the step.
python scripts/ncu_gemm_top.py [--top 8] [--time]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from klab_multimodalmodel_b200 import ops as O

BF, F32 = torch.bfloat16, torch.float32
# (M, N, K, a_mn, b_mn, out dtype, bias, act, residual dtype, aux_in dtype, aux_out, dropout p, accumulate, ldd, count per step)
TOP = [
    (6144, 4096, 1024, False, True, BF, False, 3, None, BF, False, 0.1, False, 4096, 24),
    (6144, 4096, 1024, False, False, BF, False, 1, None, None, False, 0.1, False, 4096, 24),
    (6144, 1024, 4096, False, False, BF, False, 0, BF, None, False, 0.1, False, 1024, 24),
    (16384, 2048, 512, False, True, BF, False, 6, None, BF, False, 0.0, False, 2048, 18),
    (6144, 1024, 4096, False, True, BF, False, 0, None, None, False, 0.0, False, 1024, 24),
    (1024, 1024, 2048, True, True, F32, False, 0, None, None, False, 0.0, False, 1024, 72),
    (1024, 4096, 6144, True, True, F32, False, 0, None, None, False, 0.0, False, 4096, 24),
    (16384, 2048, 512, False, False, BF, True, 5, None, None, True, 0.0, False, 2048, 18),
    (2048, 3072, 1024, False, False, BF, False, 0, None, None, False, 0.0, False, 3072, 48),
    (2048, 1024, 1024, False, True, BF, False, 0, None, None, False, 0.0, False, 1024, 72),
    (262144, 512, 128, False, True, BF, False, 6, None, BF, False, 0.0, False, 512, 2),
]

ap = argparse.ArgumentParser()
ap.add_argument("--top", type=int, default=8)
ap.add_argument("--time", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
seedp = torch.zeros(1, dtype=torch.int64, device=dev)
sets = [(s, bench._sig_operands(s[:14], dev, 1, seedp)[0]) for s in TOP[:a.top]]
for _ in range(3):
    for s, (A, B, kw) in sets:
        O.gemm(A, B, s[0], s[1], s[2], **kw)
torch.cuda.synchronize()
if a.time:
    for s, (A, B, kw) in sets:
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        for _ in range(10):
            O.gemm(A, B, s[0], s[1], s[2], **kw)
        en.record()
        torch.cuda.synchronize()
        ms = st.elapsed_time(en) / 10
        print(f"M={s[0]} N={s[1]} K={s[2]} a_mn={int(s[3])} b_mn={int(s[4])} act={s[7]}: {ms * 1e3:.1f} us  {2.0 * s[0] * s[1] * s[2] / ms / 1e9:.0f} TFLOP/s")
torch.cuda.profiler.start()
for s, (A, B, kw) in sets:
    O.gemm(A, B, s[0], s[1], s[2], **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
