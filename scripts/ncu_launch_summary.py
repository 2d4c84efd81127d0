"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`), no GPU needed.
python scripts/ncu_launch_summary.py gpurun_out/X.csv > profiles/rNN_ncu_launch_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1], errors="replace") if l.startswith('"')]
rd = csv.DictReader(rows)
tot, cnt = defaultdict(float), defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"(klab::)?(<unnamed>|\(anonymous namespace\))::", "", name).replace("klab::", "")
    m = re.match(r"([\w:]+(?:<[^()]*>)?)", name)
    name = m.group(1) if m else name
    name = name.replace("(bool)", "")
    if name.startswith("at::") or "at::native" in name:
        name = "torch:" + (re.search(r"(\w+_kernel)", name).group(1) if re.search(r"(\w+_kernel)", name) else name[:40])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"# {sum(cnt.values())} launches, sum of kernel time {total / 1e3:.2f} ms")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k:98s} {cnt[k]:5d}  {tot[k] / 1e3:8.3f} ms  {100 * tot[k] / total:4.1f}%")
