"""Top stall locations of one kernel from an .ncu-rep (SASS view with source correlation when -lineinfo is present).
python scripts/ncu_source_hot.py rep kernel_regex [--top 30] [--view sass|source]"""
import argparse, csv, io, subprocess, collections
ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("kernel"); ap.add_argument("--top", type=int, default=30)
ap.add_argument("--cuda-source", action="store_true")
a = ap.parse_args()
cmd = ["ncu", "-i", a.rep, "--page", "source", "--csv", "--kernel-name", "regex:" + a.kernel]
if a.cuda_source:
    cmd += ["--print-source", "cuda,sass"]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] in ("Address", "#", "Line"))
names = rows[hdr]
print("columns:", names[:8])
si = names.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in names else None
srci = names.index("Source")
data = []
for r in rows[hdr + 1:]:
    if len(r) != len(names) or not r[si].strip().isdigit():
        if data:
            break                      # next kernel instance: keep the first only
        continue
    data.append(r)
tot = sum(int(r[si] or 0) for r in data)
print("total samples", tot)
stall_cols = [i for i, n in enumerate(names) if n.startswith("stall_") or n.lower().startswith("warp stall")]
order = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:a.top]
for i in sorted(order):
    r = data[i]
    extra = ""
    named = [(names[j], r[j]) for j in range(len(names)) if names[j].startswith("stall_") and r[j] not in ("0", "")]
    named.sort(key=lambda kv: -int(kv[1]))
    extra = " ".join(f"{k[6:]}={v}" for k, v in named[:3])
    print(f"{i:5d} {100.0 * int(r[si] or 0) / max(tot, 1):5.1f}%  {r[srci].strip()[:90]:90s} {extra}")
