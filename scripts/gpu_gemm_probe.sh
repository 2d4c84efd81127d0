#!/bin/bash
# Runs every GEMM probe variant in its own process (a device trap must not poison the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/probe.log
for v in kk_64 kk_128 kk_256 kk_ragged kn_256 kn_ragged nn_256 nn_ragged nk_128; do
  timeout 180 python scripts/gpu_gemm_probe.py $v 2>&1 | tail -8 | tee -a gpurun_out/probe.log
done
timeout 300 python scripts/gpu_gemm_probe.py perf 2>&1 | tail -30 | tee -a gpurun_out/probe.log
