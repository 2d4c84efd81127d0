"""Development aid: per-kernel GPU time of one data-parallel training step on rank 0 (torch.profiler / CUPTI), with the NCCL
kernels and the idle time at the end of backward called out.
torchrun --nproc-per-node N scripts/profile_ddp.py"""
import os
import re
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_MAX_CTAS", "16")
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP
from torch.profiler import ProfilerActivity, profile

import bench

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
w = dict(bench.WORKLOADS["2a"])
model, tcfg = bench.build_model(w, dev, "bf16")
model.transformer.train()
net = DDP(model, device_ids=[local])
from klab_multimodalmodel_b200.optim import Adam
opt = Adam(model.transformer.parameters(), lr=1e-4)
px, src, tgt = [t.to(dev) for t in bench.synth_batch(w, tcfg.vocab_size, 1234 + rank, pin=False)]


def step():
    loss = net({"pixel_values": px}, {"input_ids": src}, {"input_ids": tgt})
    lv = loss.item()
    loss.backward()
    opt.step()
    opt.zero_grad()
    return lv


for _ in range(4):
    step()
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
for _ in range(3):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0]
    nccl = [e for e in evs if "nccl" in e.name.lower()]
    ours = [e for e in evs if "nccl" not in e.name.lower()]
    t_begin = min(e.time_range.start for e in evs)
    t_end = max(e.time_range.end for e in evs)
    print(f"world {dist.get_world_size()}: wall {wall * 1e3:.1f} ms/step; profiled span {(t_end - t_begin) / 1e3:.1f} ms")
    print(f"nccl kernels: {len(nccl)}, total {sum(e.device_time for e in nccl) / 1e3:.2f} ms, "
          f"first starts at +{(min(e.time_range.start for e in nccl) - t_begin) / 1e3:.1f} ms, last ends at +{(max(e.time_range.end for e in nccl) - t_begin) / 1e3:.1f} ms")
    last_compute_before_adam = [e for e in ours if "adam" in e.name]
    if last_compute_before_adam:
        a = last_compute_before_adam[0]
        prev = max((e.time_range.end for e in ours if e.time_range.end <= a.time_range.start), default=a.time_range.start)
        print(f"adam starts at +{(a.time_range.start - t_begin) / 1e3:.1f} ms; idle gap before it {(a.time_range.start - prev) / 1e3:.2f} ms")
    # GPU occupancy of the step: union of the intervals of our kernels (any stream) = busy; the rest of the span = idle
    iv = sorted((e.time_range.start, e.time_range.end) for e in ours)
    busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    gaps = []
    for s_, e_ in iv[1:]:
        if s_ > cur_e:
            busy += cur_e - cur_s
            gaps.append((s_ - cur_e, cur_e - t_begin))
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    busy += cur_e - cur_s
    print(f"our kernels: busy (union) {busy / 1e3:.2f} ms, idle inside the span {(t_end - t_begin - busy) / 1e3:.2f} ms over {len(gaps)} gaps; "
          f"sum of kernel durations {sum(e.device_time for e in ours) / 1e3:.2f} ms")
    big = sorted(gaps, reverse=True)[:8]
    print("largest gaps (ms @ offset ms): " + ", ".join(f"{g / 1e3:.2f}@{o / 1e3:.1f}" for g, o in big))
    agg = {}
    for e in evs:
        k = re.sub(r"\(.*", "", e.name).replace("void ", "").replace("klab::(anonymous namespace)::", "")[:60]
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + e.device_time)
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"{k:60s} {c:5d} {t / 1e3:8.2f} ms")
    red = model._klab_reducer
    print("buckets", red.buckets_last_backward if red else None)
dist.barrier()
dist.destroy_process_group()
