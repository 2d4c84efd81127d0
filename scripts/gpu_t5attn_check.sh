#!/bin/bash
# Development aid: parity + timing of the two-CTA-per-SM T5 attention kernels against the 512-thread ones (one gpurun call).
#   bash scripts/gpu_t5attn_check.sh <tag> [bench]
set -u
out=gpurun_out
tag=${1:-r2z}
for v in 1 0; do
  KLAB_T5_ATTN_2CTA=$v timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q -k "t5_attention and not decode" > $out/${tag}_attn_tests_2cta$v.log 2>&1; echo "attn tests 2cta=$v rc=$?"; tail -1 $out/${tag}_attn_tests_2cta$v.log
  KLAB_T5_ATTN_2CTA=$v timeout 200 python scripts/attn_probe.py --time > $out/${tag}_probe_2cta$v.txt 2>&1; echo "probe 2cta=$v rc=$?"
done
paste <(grep "t5" $out/${tag}_probe_2cta1.txt) <(grep "t5" $out/${tag}_probe_2cta0.txt | awk '{print $(NF-1)}')
timeout 600 python -m pytest tests/test_step_parity_gpu.py -x -q > $out/${tag}_step_tests.log 2>&1; echo "step tests rc=$?"; tail -2 $out/${tag}_step_tests.log
if [ "${2:-}" = bench ]; then
for v in 1 0; do
  KLAB_T5_ATTN_2CTA=$v timeout 400 python bench.py --no-extras --no-cpu-baseline --steps 10 > $out/${tag}_bench_2cta$v.json 2> $out/${tag}_bench_2cta$v.err; echo "bench 2cta=$v rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/${tag}_bench_2cta$v.json") if l.startswith("{")][0])
    print("2cta=$v", d["ms_per_step"], d["ms_per_step_isolated"], d["e2e"]["value"], d["roofline"]["frac"], d["loss"])
except Exception as e:
    print("no bench line", e)
PY
done
fi
