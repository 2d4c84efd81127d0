"""Development aid: the four attention kernels at their workload-2a geometries, once each inside a cudaProfilerStart/Stop range
(for `ncu --profile-from-start off --set full`), plus CUDA-event timings of every Swin stage / T5 site.
python scripts/attn_probe.py [--time]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from klab_multimodalmodel_b200 import ops as O

ap = argparse.ArgumentParser()
ap.add_argument("--time", action="store_true")
ap.add_argument("--stage", type=int, default=2)
ap.add_argument("--big", action="store_true", help="profile the 384^2 / 12x12-window Swin kernels and the 176-token T5 kernels instead")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
B = 64


def swin_case(stage, big=False):
    res, C, heads = [(64, 128, 4), (32, 256, 8), (16, 512, 16), (8, 1024, 32)][stage]
    w, shift, hd = 8, (4 if res > 8 else 0), 32
    if big:                                                    # 384^2 inputs, 12 x 12 windows (bench workload 4a)
        res, w, shift = res * 3 // 2, 12, (6 if res > 8 else 0)
    T = B * res * res
    qkv = torch.randn(T, 3 * C, device=dev).bfloat16()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    ls = torch.full((heads,), 2.3, device=dev)
    bias = torch.randn(heads, w * w, w * w, device=dev)
    ctx, lse = O.swin_attention_fwd(q, k, v, B, res, heads, hd, w, shift, ls, bias)
    dctx = torch.randn_like(ctx)
    dqkv = torch.empty_like(qkv)
    fwd = lambda: O.swin_attention_fwd(q, k, v, B, res, heads, hd, w, shift, ls, bias)
    bwd = lambda: O.swin_attention_bwd(q, k, v, ctx, dctx, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], B, res, heads, hd, w, shift, ls, bias, lse)
    return fwd, bwd


def t5_case(Lq, Lk, causal, has_bias):
    H, dk = 16, 64
    inner = H * dk
    qb = torch.randn(B * Lq, inner, device=dev).bfloat16() * 0.3
    kvb = torch.randn(B * Lk, 2 * inner, device=dev).bfloat16() * 0.3
    q, k, v = qb, kvb[:, :inner], kvb[:, inner:]
    table = torch.randn(32, H, device=dev) if has_bias else None
    lut, rz = (None, 0)
    if has_bias:
        lut, rz = O.t5_rel_bucket_lut(Lq, Lk, not causal, 32, 128)
        lut = lut.to(dev)
    seedp = torch.zeros(1, dtype=torch.int64, device=dev)
    kw = dict(bias_table=table, lut=lut, rel_zero=rz, causal=causal, dropout_p=0.1, seed=3, seed_ptr=seedp)
    ctx, lse = O.t5_attention_fwd(q, k, v, B, H, Lq, Lk, dk, **kw)
    dctx = torch.randn_like(ctx)
    dq, dkv = torch.empty_like(qb), torch.empty_like(kvb)
    dtab = torch.zeros_like(table) if has_bias else None
    fwd = lambda: O.t5_attention_fwd(q, k, v, B, H, Lq, Lk, dk, **kw)
    bwd = lambda: O.t5_attention_bwd(q, k, v, ctx, dctx, lse, dq, dkv[:, :inner], dkv[:, inner:], B, H, Lq, Lk, dk, dbias_table=dtab, **kw)
    return fwd, bwd


def timed(name, fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    print(f"{name:28s} {s.elapsed_time(e) / iters * 1e3:8.1f} us")


if a.time:
    for st in range(4):
        f, b = swin_case(st)
        timed(f"swin stage {st + 1} fwd", f)
        timed(f"swin stage {st + 1} bwd", b)
    for st in range(4):
        f, b = swin_case(st, big=True)
        timed(f"swin 384/w12 stage {st + 1} fwd", f)
        timed(f"swin 384/w12 stage {st + 1} bwd", b)
    for name, (Lq, Lk, causal, hb) in {"t5 enc self 176x176": (176, 176, False, True), "t5 dec self 128x128 causal": (128, 128, True, True),
                                       "t5 cross 128x176": (128, 176, False, False)}.items():
        f, b = t5_case(Lq, Lk, causal, hb)
        timed(name + " fwd", f)
        timed(name + " bwd", b)
    for name, (Lq, Lk, causal, hb) in {"t5 enc self 96x96": (96, 96, False, True), "t5 frozen enc 32x32": (32, 32, False, True),
                                       "t5 dec self 32x32 causal": (32, 32, True, True), "t5 cross 32x96": (32, 96, False, False)}.items():
        f, b = t5_case(Lq, Lk, causal, hb)
        timed(name + " fwd", f)
        timed(name + " bwd", b)
else:
    sf, sb = swin_case(a.stage, big=a.big)
    tf, tb = t5_case(176, 176, False, True) if a.big else t5_case(96, 96, False, True)
    for fn in (sf, sb, tf, tb):
        fn(); fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for fn in (sf, sb, tf, tb):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
