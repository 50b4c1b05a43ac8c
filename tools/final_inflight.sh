#!/bin/bash
# Refresh of the headline evidence after the in-flight / throughput-sampling change:
#   gpurun --timeout 900 -- 'bash tools/final_inflight.sh'
set -x
cd "$(dirname "$0")/.."
python bench.py --steps 100 --warmup 5 > gpurun_out/final2_bench_1gpu.json 2> gpurun_out/final2_bench_1gpu.err
python bench.py --steps 100 --warmup 5 --in-flight 1 --no-cpu-baseline > gpurun_out/final2_bench_1gpu_serial.json 2> gpurun_out/final2_bench_1gpu_serial.err
python bench.py --workload detector --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/final2_bench_detector.json 2> gpurun_out/final2_bench_detector.err
python tools/overlap_probe.py > gpurun_out/final2_overlap_probe.json 2> gpurun_out/final2_overlap_probe.err
# ncu: launch list of the bench command (eager launches; serialised, cold: shares only) ...
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/final2_launches_bench.log 2>&1
# ... and the counters of the throughput variant of the sampling kernel
LEAN=1 ncu --set full --import-source on --clock-control none -k regex:'fps_sorted' -c 1 -f -o /tmp/lean \
    python tools/run_ops.py > gpurun_out/final2_lean_run.log 2>&1
ncu -i /tmp/lean.ncu-rep --page raw --csv > /tmp/lean_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/lean_raw.csv > gpurun_out/final2_lean_ncu.json
ls -la gpurun_out/final2_*
