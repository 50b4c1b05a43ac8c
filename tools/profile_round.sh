#!/bin/bash
# One gpurun call that refreshes the ncu evidence under gpurun_out/ (copied into profiles/ by hand):
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh'
# 1. launch list of the bench command (serialised, cold: shares only)
# 2. ncu --set full of every distinct kernel of one eager backbone forward
set -x
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'fps_|ball_query|grid_|sa_mlp|fp_mlp' -s 36 -c 36 -f -o /tmp/kernels \
    python tools/run_ops.py > gpurun_out/kernels_run.log 2>&1
ncu -i /tmp/kernels.ncu-rep --page raw --csv > /tmp/kernels_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/kernels_raw.csv > gpurun_out/kernels_ncu.json
ls -la gpurun_out/kernels* gpurun_out/launches*
