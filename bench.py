#!/usr/bin/env python
"""Headline benchmark: train samples/sec of the Swin-V2 + T5 image-caption step (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 2a|2b|1a|tiny] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is exactly what /root/reference/train.py:54-71 does per batch with accumulation_steps = 1:
    (host->device copy of the batch) -> loss = model(images, src, tgt) -> loss.item() -> loss.backward() ->
    optimizer.step() (Adam over model.transformer.parameters(), train.py:28) -> optimizer.zero_grad()
with `model.transformer.train()` (T5 dropout p = 0.1 active, train.py:52) and DistributedDataParallel for N > 1.

Workload at N = 1 ("2a", BASELINE.json configs[1] realised as SURVEY.md 8d recommends because Swin-B + T5-base cannot be
concatenated, 1024 != 768): Swin-B/256/window 8 trained jointly + T5-large, bf16 tensor-core path, batch 64 per GPU,
32 source + 32 target tokens, synthetic inputs, random-init weights.  One JSON line on stdout (rank 0).

  value          : whole-job samples/s with the batch already resident in HBM (timed with CUDA events, max over ranks)
  e2e            : same through the public API with HOST (pinned) inputs: H2D copies and the loss read-back inside the timed region
  roofline       : the dominant kernel family (tcgen05 GEMM, ~97% of the FLOPs): algorithmic FLOPs of every GEMM launch of one
                   step / CUDA-event duration of those launches, against MEASURED_PEAKS.json (sustained bf16 figure)
  cpu_baseline   : the oracle port of the reference path (oracle/, plain torch fp32) on the host cores, bounded sample
  --impl reference : that same CPU path as its own arm (the reference is Python glue over `transformers`; it has no build to
                   compile, /root/reference does not exist on the GPU box, so the oracle port stands in: kind "port")
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# GFLOP per sample, fwd+bwd, counted on the reference graph
WORKLOADS = {
    "2a": dict(desc="MSCOCO caption train step, Swin-B/256/w8 trained jointly + T5-large (runnable realisation of configs[1])",
               swin=dict(image_size=256, embed_dim=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32), window_size=8),
               t5="t5-large", batch=64, l_src=32, l_tgt=32, train_swin=True, gflop_per_sample=422.14),
    "2b": dict(desc="MSCOCO caption train step, Swin-S/256/w8 trained jointly + T5-base",
               swin=dict(image_size=256, embed_dim=96, depths=(2, 2, 18, 2), num_heads=(3, 6, 12, 24), window_size=8),
               t5="t5-base", batch=64, l_src=32, l_tgt=32, train_swin=True, gflop_per_sample=157.01),
    "1a": dict(desc="MSCOCO caption train step, frozen Swin-T/224/w7 + T5-base (runnable realisation of configs[0])",
               swin=dict(image_size=224, embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=7),
               t5="t5-base", batch=2, l_src=32, l_tgt=32, train_swin=False, gflop_per_sample=87.53),
    "tiny": dict(desc="smoke-sized step", swin=dict(image_size=64, embed_dim=32, depths=(2, 2, 2), num_heads=(1, 2, 4), window_size=4),
                 t5=dict(vocab_size=512, d_model=128, d_ff=256, num_layers=2, num_heads=2), batch=4, l_src=8, l_tgt=8,
                 train_swin=True, gflop_per_sample=0.05),
}
T5_NAMED = {"t5-small": dict(d_model=512, d_ff=2048, num_layers=6, num_heads=8),
            "t5-base": dict(d_model=768, d_ff=3072, num_layers=12, num_heads=12),
            "t5-large": dict(d_model=1024, d_ff=4096, num_layers=24, num_heads=16)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(tflops=p.get("bf16_tflops_sustained", 1378.2), burst=p.get("bf16_tflops", 1632.1), hbm=p.get("hbm_gbs", 6450.3),
                    source="measured")
    return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback")


def synth_batch(w, vocab, seed, pin):
    g = torch.Generator().manual_seed(seed)
    s = w["swin"]["image_size"]
    px = torch.randn(w["batch"], 3, s, s, generator=g)
    src = torch.randint(2, vocab - 28, (w["batch"], w["l_src"]), generator=g)
    tgt = torch.randint(2, vocab - 28, (w["batch"], w["l_tgt"]), generator=g)
    tgt[:, -1] = 1
    if pin:
        px, src, tgt = px.pin_memory(), src.pin_memory(), tgt.pin_memory()
    return px, src, tgt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        return len(self.rows)

    def summary(self, lo, hi):
        rows = [r for r in self.rows[lo:max(hi, lo + 1)] if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons, "samples": len(rows)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path
# --------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(w, steps, warmup, sample_batch, log=lambda *_: None):
    """fwd + bwd + Adam step of the reference path restated in oracle/ (plain torch fp32 on the host cores)."""
    from oracle.caption_model import caption_loss, swin_param_shapes, t5_param_shapes
    from oracle.swinv2 import SwinDims
    from oracle.t5 import T5Dims
    torch.set_num_threads(os.cpu_count() or 1)
    t5kw = T5_NAMED[w["t5"]] if isinstance(w["t5"], str) else w["t5"]
    t5 = T5Dims(**t5kw)
    sw = dict(w["swin"])
    sw.setdefault("pretrained_window_sizes", (0,) * len(sw["depths"]))
    swin = SwinDims(**sw)
    g = torch.Generator().manual_seed(0)

    def make(shapes, train):
        sd, tied = {}, ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "lm_head.weight")
        for k, shp in shapes.items():
            if k in tied:
                continue
            if len(shp) == 1 and k.endswith("weight"):
                t = torch.ones(shp)
            elif k.endswith("bias"):
                t = torch.zeros(shp)
            elif k.endswith("logit_scale"):
                t = torch.full(shp, 2.302585)
            else:
                t = torch.randn(shp, generator=g) * (1.0 if k == "shared.weight" else shp[-1] ** -0.5 * 0.5)
            sd[k] = t.requires_grad_(train)
        for k in tied:
            if k in shapes:
                sd[k] = sd["shared.weight"]
        return sd

    sds = {"language_model": make(t5_param_shapes(t5, encoder_only=True), False),
           "image_model": make(swin_param_shapes(swin), w["train_swin"]),
           "transformer": make(t5_param_shapes(t5), True)}
    params = list({id(v): v for v in sds["transformer"].values()}.values())
    opt = torch.optim.Adam(params, lr=1e-3)
    wb = dict(w, batch=sample_batch)
    px, src, tgt = synth_batch(wb, t5.vocab_size, 1234, pin=False)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = caption_loss(px, src, tgt, sds, t5, swin, t5)
        lv = loss.item()
        loss.backward()
        opt.step()
        opt.zero_grad()
        dt = time.perf_counter() - t0
        log(f"cpu step {i}: {dt:.2f}s loss {lv:.3f}")
        if i >= warmup:
            times.append(dt)
    return times


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample_batch
    times = cpu_reference_steps(w, args.steps, args.warmup, sample, log=lambda m: print(m, file=sys.stderr))
    ms = 1e3 * sum(times) / len(times)
    val = sample / (ms / 1e3)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "train samples/sec (Swin+T5 caption step)", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "name": args.workload, "per_step_sample_batch": sample, "l_src": w["l_src"], "l_tgt": w["l_tgt"]},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} samples/step of the same model + sequence lengths, fwd+bwd+Adam, torch fp32, {cores} threads"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def build_model(w, device, dtype):
    import types

    from klab_multimodalmodel_b200.modeling import Swinv2Config, T5Config, init_swin_, init_t5_
    from klab_multimodalmodel_b200.models.model import MyModel
    t5kw = T5_NAMED[w["t5"]] if isinstance(w["t5"], str) else w["t5"]
    tcfg = T5Config(**t5kw)
    sw = dict(w["swin"])
    sw.setdefault("pretrained_window_sizes", (0,) * len(sw["depths"]))
    scfg = Swinv2Config(**sw)
    args = types.SimpleNamespace(result_dir="/tmp", language_model_name=tcfg, image_model_name=scfg, image_model_train=w["train_swin"],
                                 transformer_model_name=tcfg, compute_dtype=dtype)
    model = MyModel(args)
    init_t5_(model.language_model, seed=1)
    init_swin_(model.image_model, seed=2)
    init_t5_(model.transformer, seed=3)
    return model.to(device), tcfg


def gemm_roofline(model_step, pk, verbose=False, check=False):
    """Roofline of the dominant kernel family (tcgen05 GEMM): (1) run one step eagerly and record the signature of every GEMM
    launch (M, N, K, operand majors, output dtype AND the fused epilogue: bias / activation / residual / aux tensors / dropout /
    accumulate); (2) launch each distinct signature back to back on the current stream with that same epilogue, rotating over
    operand copies that together exceed the 126 MB L2, timed with CUDA events; (3) achieved = sum of algorithmic FLOPs (2 M N K)
    of the step's launches / sum(count x measured duration)."""
    import collections

    from klab_multimodalmodel_b200 import ops as O
    from klab_multimodalmodel_b200.graphs import POOL
    sigs = collections.Counter()
    orig = O.gemm

    def rec(a, b, M, N, K, **kw):
        out = orig(a, b, M, N, K, **kw)
        if a.dtype == torch.bfloat16:
            res, ain, aout = kw.get("residual"), kw.get("aux_in"), kw.get("aux_out")
            sigs[(M, N, K, bool(kw.get("a_mn")), bool(kw.get("b_mn")), out.dtype, kw.get("bias") is not None, int(kw.get("act", 0)),
                  None if res is None else res.dtype, None if ain is None else ain.dtype, aout is not None,
                  float(kw.get("dropout_p", 0.0)), bool(kw.get("accumulate", False)), out.stride(0))] += 1
        return out

    was = POOL.enabled
    O.gemm, POOL.enabled = rec, False
    try:
        model_step()
        torch.cuda.synchronize()
    finally:
        O.gemm, POOL.enabled = orig, was
    tot_ms, tot_fl, rows = 0.0, 0.0, []
    dev = torch.device("cuda", torch.cuda.current_device())
    seedp = torch.zeros(1, dtype=torch.int64, device=dev)
    for sig, cnt in sigs.items():
        M, N, K, a_mn, b_mn, od, has_bias, act, res_dt, ain_dt, has_aout, p_drop, acc, ldd = sig
        esz = 2 if od == torch.bfloat16 else 4
        per = (M * K + N * K) * 2 + M * ldd * esz * (1 + (res_dt is not None) + (ain_dt is not None) + has_aout)
        copies = max(1, min(8, (300 << 20) // max(per, 1)))
        As = [torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16() for _ in range(copies)]
        Bs = [torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16() for _ in range(copies)]
        Ds = [torch.zeros(M, ldd, device=dev, dtype=od)[:, :N] for _ in range(copies)]
        kws = []
        for i in range(copies):
            kw = dict(a_mn=a_mn, b_mn=b_mn, out=Ds[i], act=act, dropout_p=p_drop, accumulate=acc, seed=5, seed_ptr=seedp if p_drop > 0 else None)
            if has_bias:
                kw["bias"] = torch.randn(N, device=dev)
            if res_dt is not None:
                kw["residual"] = torch.randn(M, N, device=dev).to(res_dt)
            if ain_dt is not None:
                kw["aux_in"] = torch.randn(M, N, device=dev).to(ain_dt)
            if has_aout:
                kw["aux_out"] = torch.empty(M, N, device=dev, dtype=od)
            kws.append(kw)
        for i in range(3):
            orig(As[i % copies], Bs[i % copies], M, N, K, **kws[i % copies])
        iters = 8
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()                      # replay from a CUDA graph, as the step does: no host launch cost in the timing
        with torch.cuda.graph(g):
            for i in range(iters):
                orig(As[i % copies], Bs[i % copies], M, N, K, **kws[i % copies])
        g.replay()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / iters
        del g
        err = None
        if check and p_drop == 0.0 and not acc and act in (0, 1, 2):
            kw = dict(kws[0]); D = torch.zeros(M, ldd, device=dev, dtype=od)[:, :N]; kw["out"] = D
            orig(As[0], Bs[0], M, N, K, **kw)
            A32 = As[0].float().t() if a_mn else As[0].float()
            B32 = Bs[0].float() if b_mn else Bs[0].float().t()
            ref = A32 @ B32
            if has_bias:
                ref = ref + kw["bias"]
            if act == 1:
                ref = torch.relu(ref)
            elif act == 2:
                ref = torch.nn.functional.gelu(ref)
            if res_dt is not None:
                ref = ref + kw["residual"].float()
            err = ((D.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-6)).item()
        fl = 2.0 * M * N * K
        tot_ms += cnt * ms
        tot_fl += cnt * fl
        rows.append((cnt * ms, cnt, M, N, K, int(a_mn), int(b_mn), ms, fl / ms / 1e9, sig[6:13], err))
        del As, Bs, Ds, kws
    if verbose:
        for r in sorted(rows, reverse=True)[:(200 if check else 40)]:
            print("gemm sig: total %.2f ms  x%d  M=%d N=%d K=%d a_mn=%d b_mn=%d  %.1f us  %.0f TFLOP/s  epi(bias,act,res,auxin,auxout,p,acc)=%s%s" %
                  (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7] * 1e3, r[8], r[9], "" if r[10] is None else "  relerr %.2e" % r[10]),
                  file=sys.stderr)
    ach = tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
    return {"bound": "tensor", "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"], "traffic": None,
            "kernel": "gemm_bf16_tc_kernel (tcgen05)", "launches_per_step": int(sum(sigs.values())), "distinct_shapes": len(sigs),
            "gemm_ms_per_step": tot_ms, "gemm_gflop_per_step": tot_fl / 1e9, "peak_source": pk["source"] + " bf16 sustained",
            "method": "every GEMM signature of one step (shape, operand majors and fused epilogue: bias / activation / residual / "
                      "aux tensors / dropout) replayed back to back from a CUDA graph, operands rotated beyond L2, CUDA events"}


def run_ours(args, w):
    import torch.distributed as dist
    from klab_multimodalmodel_b200 import _lib as L
    from klab_multimodalmodel_b200 import ops as O
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # the collective shares the GPU with the persistent tcgen05 kernels (which hand out work dynamically and simply run
        # on the SMs that are left): cap its CTAs so that it never takes more than ~10 % of the machine
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=int(os.environ.get("KLAB_NCCL_TIMEOUT_S", "600"))))
    L.check(L.lib().klab_check_device())
    model, tcfg = build_model(w, dev, args.dtype)
    model.transformer.train()                                        # train.py:52
    net = model
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        net = DDP(model, device_ids=[local])
    if args.optimizer == "klab":                                      # N1: fused multi-tensor Adam, same semantics as train.py:28
        from klab_multimodalmodel_b200.optim import Adam
        opt = Adam(model.transformer.parameters(), lr=1e-4)
    else:
        opt = torch.optim.Adam(model.transformer.parameters(), lr=1e-4)
    px_h, src_h, tgt_h = synth_batch(w, tcfg.vocab_size, 1234 + rank, pin=True)
    px_d, src_d, tgt_d = px_h.to(dev), src_h.to(dev), tgt_h.to(dev)
    if args.gemm_probe:
        def probe_step():
            loss = model({"pixel_values": px_d}, {"input_ids": src_d}, {"input_ids": tgt_d})
            loss.backward()
            opt.zero_grad()
        probe_step()
        r = gemm_roofline(probe_step, peaks(), verbose=True, check=True)
        emit(r)
        return
    h2d = px_h.numel() * 4 + src_h.numel() * 8 + tgt_h.numel() * 8
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)    # > 126 MB L2

    wd = int(os.environ.get("KLAB_BENCH_WATCHDOG_S", "0"))          # debugging aid: dump every thread's stack if a step stalls
    if wd:
        import faulthandler

    def step(px, src, tgt, read_loss=True):
        if wd:
            faulthandler.dump_traceback_later(wd, exit=True)
        loss = net({"pixel_values": px}, {"input_ids": src}, {"input_ids": tgt})
        lv = loss.item() if read_loss else None                          # train.py:59 reads the loss every step
        loss.backward()
        opt.step()
        opt.zero_grad()
        return lv

    def resident_step():
        return step(px_d, src_d, tgt_d)

    def e2e_step():
        return step(px_h.to(dev, non_blocking=True), src_h.to(dev, non_blocking=True), tgt_h.to(dev, non_blocking=True))

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total = 0.0
        last = None
        for _ in range(k):
            flush.zero_()                                                   # L2 flush between timed iterations (untimed)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            last = fn()
            e.record()
            torch.cuda.synchronize()
            total += s.elapsed_time(e)
        if world > 1:
            dist.barrier()
        t = torch.tensor([total], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), last

    for _ in range(max(args.warmup, 3)):
        resident_step()
    torch.cuda.synchronize()
    if args.profile_range:                                # ncu --profile-from-start off: exactly the timed steps are profiled
        torch.cuda.profiler.start()
        for _ in range(args.steps):
            resident_step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3)
    lo = sampler.mark() if sampler else 0
    from klab_multimodalmodel_b200.graphs import POOL
    l0 = O.launch_count() + POOL.replayed_kernels
    ms_res, loss_v = timed(resident_step, args.steps)
    launches = O.launch_count() + POOL.replayed_kernels - l0
    hi = sampler.mark() if sampler else 0
    ms_e2e, _ = timed(e2e_step, args.steps)
    clocks = sampler.summary(lo, hi) if sampler else None
    if sampler:
        sampler.stop()
    dp_sync = None
    if world > 1:                                       # data-parallel sanity: after K averaged steps every rank holds the same weights
        probe = torch.stack([p.detach().double().sum() for p in list(model.transformer.parameters())[:8] + list(model.transformer.parameters())[-8:]])
        lo_, hi_ = probe.clone(), probe.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        dp_sync = bool(torch.equal(lo_, hi_))
    B = w["batch"]
    ms_step = ms_res / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e_val = world * B / (ms_e2e / args.steps * 1e-3)
    pk = peaks()
    if world > 1:                                       # collectives are over: what follows is rank-0-local reporting
        dist.barrier()
        dist.destroy_process_group()
        if getattr(model, "_klab_reducer", None) is not None:
            model._klab_reducer.enabled = False
    def local_step():                                   # rank-local (no DDP collectives): only rank 0 runs the roofline probe
        loss = model({"pixel_values": px_d}, {"input_ids": src_d}, {"input_ids": tgt_d})
        loss.backward()
        opt.zero_grad()

    if wd:
        faulthandler.cancel_dump_traceback_later()
    roof = gemm_roofline(local_step, pk, verbose=args.verbose) if rank == 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sb = args.cpu_sample_batch
        times = cpu_reference_steps(w, 2, 1, sb, log=lambda m: print(m, file=sys.stderr))
        cores = os.cpu_count() or 1
        cpu = {"value": sb / (sum(times) / len(times)), "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{sb} samples/step of the same model + sequence lengths (1 warm-up + 2 timed steps), fwd+bwd+Adam, torch fp32, {cores} threads"}
    if rank == 0:
        step_tflops = value * w["gflop_per_sample"] / 1e3
        line = {
            "metric": "train samples/sec (Swin+T5 caption step)", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "name": args.workload, "batch_per_gpu": B, "global_batch": B * world, "l_src": w["l_src"],
                       "l_tgt": w["l_tgt"], "parallelism": f"dp{world}", "optimizer": ("klab_multimodalmodel_b200.optim.Adam (fused multi-tensor kernel, torch.optim.Adam semantics)" if args.optimizer == "klab"
                                     else "torch.optim.Adam") + " over transformer params (train.py:28)",
                       "dropout": "T5 p=0.1 active (train.py:52)", "l2_flush": "256 MiB buffer zeroed between timed iterations"},
            "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "step_model_tflops": step_tflops, "step_frac_of_peak": step_tflops / pk["tflops"], "loss": loss_v,
            "cuda_graphs": POOL.stats(),
        }
        if world > 1:
            red = getattr(model, "_klab_reducer", None)
            line["data_parallel"] = {"weights_identical_across_ranks": dp_sync, "reducer": "klab GradReducer (grouped in-place NCCL all-reduce per bucket)" if red else "torch DDP",
                                     "buckets_per_step": getattr(red, "buckets_last_backward", None), "sm_reserve": O.sm_reserve_info(),
                                     "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS")}
        emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else written to fd 1 meanwhile (e.g. NCCL's version banner) was
    diverted to stderr by `main`."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="2a", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample-batch", type=int, default=2)
    ap.add_argument("--optimizer", default="klab", choices=["klab", "torch"],
                    help="klab = fused multi-tensor Adam (SURVEY 8f N1, same update rule); torch = stock torch.optim.Adam")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--profile-range", action="store_true",
                    help="for ncu --profile-from-start off: cudaProfilerStart/Stop around the K steps after warm-up, then exit (no JSON line)")
    ap.add_argument("--gemm-probe", action="store_true", help="development aid: only replay (and check) the step's GEMM signatures")
    args = ap.parse_args()
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                                  # libraries that print to stdout (NCCL banner) must not pollute the JSON line
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["batch"] = args.batch
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
