#!/bin/bash
# Evidence for the tf32x3 kernels (profiles/r02_tf32_*): MMA issue-rate probe, same-box A/B of the kernel variants with the
# bitwise comparison, in-kernel cycle counters, and the ncu launch list of K1 (1 x 200 frames).  Run under gpurun.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 tools/probes/mma_rate.bin > gpurun_out/r02_mma_rate.log 2>&1; tail -20 gpurun_out/r02_mma_rate.log
timeout 600 python tools/pair_ab_tf32.py 10 > gpurun_out/r02_tf32_pair_ab.log 2>&1; tail -5 gpurun_out/r02_tf32_pair_ab.log
timeout 300 python tools/layer_timing_tf32.py > gpurun_out/r02_tf32_layer_timing.log 2>&1; tail -3 gpurun_out/r02_tf32_layer_timing.log
# (under ncu the engine launches tf32_flow_kernel without the cooperative attribute: ncu rejects cooperative cluster launches)
python tools/profile_step.py 1 200 256 tf32x3 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r02_launches_k1_tf32x3.csv python tools/profile_step.py 1 200 256 tf32x3 > gpurun_out/ncu_k1.log 2>&1
tail -2 gpurun_out/ncu_k1.log
