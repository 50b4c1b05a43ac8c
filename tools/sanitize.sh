#!/bin/bash
# compute-sanitizer over the cluster / DSMEM / mbarrier / tcgen05 kernels:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# A small backbone forward (both sampling variants, cluster sizes > 1, fused SA / FP kernels) under
# memcheck, racecheck (shared-memory hazards) and synccheck; logs -> gpurun_out/sanitize_*.log
set -x
cd "$(dirname "$0")/.."
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 \
      python tools/sanitize_run.py > gpurun_out/sanitize_$tool.log 2>&1
  tail -n 6 gpurun_out/sanitize_$tool.log
done
