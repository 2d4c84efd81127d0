"""Development aid: launch ONE kernel family at its workload-2a geometry a few times (for `ncu -k regex:...` captures and quick
CUDA-event timings).   python scripts/kernel_probe.py {swin_bwd|swin_fwd|t5_bwd|t5_fwd|gemm} [--iters 5]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from klab_multimodalmodel_b200 import ops as O

ap = argparse.ArgumentParser()
ap.add_argument("what")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--stage", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
B = 64


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    print(f"{a.what}: {s.elapsed_time(e) / a.iters * 1e3:.1f} us per call")


if a.what.startswith("swin"):
    res, C, heads = [(64, 128, 4), (32, 256, 8), (16, 512, 16), (8, 1024, 32)][a.stage]
    w, shift, hd = 8, (4 if res > 8 else 0), 32
    T = B * res * res
    qkv = torch.randn(T, 3 * C, device=dev).bfloat16()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    ls = torch.full((heads,), 2.3, device=dev)
    bias = torch.randn(heads, 64, 64, device=dev)
    ctx, lse = O.swin_attention_fwd(q, k, v, B, res, heads, hd, w, shift, ls, bias)
    dctx = torch.randn_like(ctx)
    dqkv = torch.empty_like(qkv)
    if a.what == "swin_fwd":
        timed(lambda: O.swin_attention_fwd(q, k, v, B, res, heads, hd, w, shift, ls, bias))
    else:
        timed(lambda: O.swin_attention_bwd(q, k, v, ctx, dctx, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], B, res, heads, hd, w, shift, ls, bias, lse))
elif a.what.startswith("t5"):
    H, L, dk = 16, 96, 64
    inner = H * dk
    qkv = torch.randn(B * L, 3 * inner, device=dev).bfloat16() * 0.3
    q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
    table = torch.randn(32, H, device=dev)
    lut, rz = O.t5_rel_bucket_lut(L, L, True, 32, 128)
    lut = lut.to(dev)
    seedp = torch.zeros(1, dtype=torch.int64, device=dev)
    ctx, lse = O.t5_attention_fwd(q, k, v, B, H, L, L, dk, bias_table=table, lut=lut, rel_zero=rz, dropout_p=0.1, seed=3, seed_ptr=seedp)
    dctx = torch.randn_like(ctx)
    dqkv = torch.empty_like(qkv)
    dtab = torch.zeros_like(table)
    if a.what == "t5_fwd":
        timed(lambda: O.t5_attention_fwd(q, k, v, B, H, L, L, dk, bias_table=table, lut=lut, rel_zero=rz, dropout_p=0.1, seed=3, seed_ptr=seedp))
    else:
        timed(lambda: O.t5_attention_bwd(q, k, v, ctx, dctx, lse, dqkv[:, :inner], dqkv[:, inner:2 * inner], dqkv[:, 2 * inner:], B, H, L, L, dk,
                                         bias_table=table, lut=lut, rel_zero=rz, dbias_table=dtab, dropout_p=0.1, seed=3, seed_ptr=seedp))
elif a.what == "gemm":
    M, N, K = 6144, 4096, 1024
    A = torch.randn(M, K, device=dev).bfloat16()
    Bm = torch.randn(N, K, device=dev).bfloat16()
    D = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    timed(lambda: O.gemm(A, Bm, M, N, K, out=D))
else:
    raise SystemExit("unknown kernel family")
