"""Per-kernel GPU time of the greedy-decode workload (bench workload 5: Swin-S/256 + T5-base, batch 256, 20 new tokens) via CUPTI
(torch.profiler); development aid.   python scripts/profile_decode.py [--batch 256] [--top 30]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from scripts.profile_step import short

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--top", type=int, default=30)
a = ap.parse_args()
w = dict(bench.WORKLOADS["5"], batch=a.batch)
dev = torch.device("cuda", 0)
model, tcfg = bench.build_model(w, dev, "bf16")
model.eval()
model.transformer.config.eos_token_id = -1
px, src, _ = [t.to(dev) for t in bench.synth_batch(w, tcfg.vocab_size, 1234, pin=False)]
from klab_multimodalmodel_b200.generation import greedy_generate
with torch.no_grad():
    emb, B, Le = model._concat_embeddings({"pixel_values": px}, {"input_ids": src})
    for _ in range(3):
        greedy_generate(model.transformer, emb, B, Le)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        greedy_generate(model.transformer, emb, B, Le)
        torch.cuda.synchronize()
agg = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time > 0:
        k = short(ev.name)
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + ev.device_time)
tot = sum(t for _, t in agg.values())
print(f"decode loop (after the encoder side): sum of kernel time {tot / 1e3:.2f} ms over {sum(c for c, _ in agg.values())} launches")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
    print(f"{k:70s} {c:6d} {t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  {t / c:7.1f} us each")
