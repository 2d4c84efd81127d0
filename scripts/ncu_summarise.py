"""Read an .ncu-rep here (no GPU needed) and print / save the per-launch metrics the roofline discussion uses.
python scripts/ncu_summarise.py gpurun_out/x.ncu-rep [--json out.json] [--sigs gemm_top]"""
import argparse
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu.sum", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__cycles_active.avg"]


def to_num(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3,
             "nsecond": 1e-9, "second": 1.0}.get(unit)
    return x * scale if scale else x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--json")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = {"kernel": r[names.index("Kernel Name")][:90]}
        for w in WANT:
            if w in names:
                i = names.index(w)
                d[w] = to_num(r[i], units[i])
        out.append(d)
    for d in out:
        t = d.get("gpu__time_duration.sum", 0.0)
        tr = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        print(f"{d['kernel'][:70]:70s} {t * 1e6:9.1f} us  dram {tr / 1e6:9.1f} MB ({tr / max(t, 1e-12) / 1e9:7.0f} GB/s)  "
              f"tensor {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', float('nan')):5.1f}%  "
              f"issue {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', float('nan')):5.1f}%  warps {d.get('sm__warps_active.avg.pct_of_peak_sustained_active', float('nan')):5.1f}%  "
              f"regs {d.get('launch__registers_per_thread', 0):.0f}  grid {d.get('launch__grid_size', 0):.0f}x{d.get('launch__block_size', 0):.0f}")
    if a.json:
        with open(a.json, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
