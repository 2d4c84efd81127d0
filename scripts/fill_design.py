"""Fill the @PLACEHOLDER@ numbers of DESIGN.md section 10 from the committed bench lines (profiles/r02_bench_final.json, _n2, _n8).
python scripts/fill_design.py"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_final.json")))
n2 = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n2.json")))
n8 = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n8.json")))
r, w = d["roofline"], d["workloads"]
v = {
    "V2A": f"{d['value']:.0f}", "MS2A": f"{d['ms_per_step']:.1f}", "E2E2A": f"{d['e2e']['value']:.0f}", "E2EPCT": f"{100 * d['e2e']['value'] / d['value']:.1f} %",
    "TORCHOPT": f"{d['torch_optimizer']['value']:.0f}", "HF2A": f"{d['hf_eager_gpu']['value']:.0f}", "XHF": f"{d['value'] / d['hf_eager_gpu']['value']:.1f}",
    "CPU": f"{d['cpu_baseline']['value']:.2f}", "GEMMSUS": f"{r['achieved']:.0f}", "FRACSUS": f"{r['frac']:.3f}", "GEMMMS": f"{r['gemm_ms_per_step']:.1f}",
    "GEMMISO": f"{r['achieved_isolated']:.0f}", "FRACBURST": f"{r['frac_burst']:.3f}", "STEPTF": f"{d['step_model_tflops']:.0f}",
    "STEPFRAC": f"{d['step_frac_of_peak']:.2f}",
    "N2V": f"{n2['value']:.0f}", "N2MS": f"{n2['ms_per_step']:.1f}", "N2EFF": f"{n2['value'] / 2 / d['value']:.3f}",
    "N8V": f"{n8['value']:.0f}", "N8MS": f"{n8['ms_per_step']:.1f}", "N8EFF": f"{n8['value'] / 8 / d['value']:.3f}",
    "V3": f"{w['3']['value']:.0f}", "MS3": f"{w['3']['ms_per_step']:.1f}", "F3": f"{w['3']['roofline']['frac']:.2f}",
    "V4A": f"{w['4a']['value']:.0f}", "MS4A": f"{w['4a']['ms_per_step']:.1f}", "F4A": f"{w['4a']['roofline']['frac']:.2f}",
    "V5": f"{w['5']['value'] / 1e3:.1f} k", "MS5": f"{w['5']['ms_per_step']:.1f}", "LOOP5": f"{w['5']['decode_loop_ms']:.1f}",
    "LOOPTOK": f"{w['5']['decode_loop_tokens_per_s'] / 1e3:.0f} k", "HF5": f"{(w['5'].get('hf_eager_gpu') or {}).get('value', float('nan')) / 1e3:.1f} k",
    "F5": f"{w['5']['roofline']['frac']:.2f}",
}
path = os.path.join(ROOT, "DESIGN.md")
tmpl = os.path.join(ROOT, "scripts", "DESIGN_section10.tmpl")
s = open(path).read()
if not os.path.exists(tmpl):                      # first run: keep the template so that later runs can refresh the numbers
    a = s.index("## 10. Results")
    open(tmpl, "w").write(s[a:])
body = open(tmpl).read()
missing = set(re.findall(r"@([A-Z0-9]+)@", body)) - set(v)
assert not missing, missing
for k, val in v.items():
    body = body.replace(f"@{k}@", val)
s = s[:s.index("## 10. Results")] + body
open(path, "w").write(s)
print("filled", len(v), "values")
