"""Per-kernel GPU time of one caption training step (CUPTI via torch.profiler; development aid -- the numbers quoted in
profiles/ come from ncu).   python scripts/profile_step.py [--workload 2a] [--batch 64] [--top 40]"""
import argparse
import os
import re
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench


def short(name: str) -> str:
    n = name.replace("klab::<unnamed>::", "").replace("klab::(anonymous namespace)::", "").replace("void ", "")
    m = re.match(r"([\w:]+)(<[^(]*>)?", n)
    if not m:
        return n[:70]
    base, targs = m.group(1), (m.group(2) or "")
    if base.startswith("at::"):
        f = re.search(r"(\w+Functor|multi_tensor_apply_kernel|\w+_kernel)", n)
        return "torch:" + (f.group(1) if f else base)
    return (base + targs.replace("__nv_bfloat16", "bf16").replace("(bool)", ""))[:70]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="2a")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    w = dict(bench.WORKLOADS[a.workload])
    if a.batch:
        w["batch"] = a.batch
    dev = torch.device("cuda", 0)
    model, tcfg = bench.build_model(w, dev, "bf16")
    model.transformer.train()
    from klab_multimodalmodel_b200.optim import Adam
    opt = Adam(model.transformer.parameters(), lr=1e-4)
    px, src, tgt = [t.to(dev) for t in bench.synth_batch(w, tcfg.vocab_size, 1234, pin=False)]

    def step():
        loss = model({"pixel_values": px}, {"input_ids": src}, {"input_ids": tgt})
        loss.backward()
        opt.step()
        opt.zero_grad()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    t_cpu = time.perf_counter() - t0          # host time to ENQUEUE one step (no sync inside)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    agg = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time > 0:
            k = short(ev.name)
            c, t = agg.get(k, (0, 0.0))
            agg[k] = (c + 1, t + ev.device_time)
    tot = sum(t for _, t in agg.values())
    print(f"workload {a.workload} batch {w['batch']}: wall {t_wall * 1e3:.1f} ms/step, host enqueue {t_cpu * 1e3:.1f} ms/step, "
          f"sum of kernel time {tot / 1e3:.1f} ms over {sum(c for c, _ in agg.values())} launches")
    print("adam table rebuilds:", getattr(opt, "rebuilds", None))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print(f"{k:70s} {c:6d} {t / 1e3:9.2f} ms {100 * t / tot:5.1f}%")


if __name__ == "__main__":
    main()
