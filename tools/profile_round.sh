#!/bin/bash
# One gpurun call that refreshes the ncu evidence under gpurun_out/ (copied into profiles/ by hand):
#   gpurun --timeout 1700 -- 'bash tools/profile_round.sh r2'
# 1. launch list of the bench command (eager issue; serialised by ncu and cold: shares only)
# 2. ncu --set full of every distinct kernel of one eager backbone forward (throughput sampling variant)
# 3. ncu --set full of the distinct conv / wgrad / BatchNorm kernels of one training step
#   STAGES="launches kernels" bash tools/profile_round.sh r2      # skip the (14-minute) training capture
set -x
R=${1:-r2}
STAGES=${STAGES:-launches kernels train}
cd "$(dirname "$0")/.."
case " $STAGES " in *" launches "*)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-ref-ext --no-train > gpurun_out/${R}_launches_bench.log 2>&1
;; esac
case " $STAGES " in *" kernels "*)
LEAN=1 ncu --set full --import-source on --clock-control none \
    -k regex:'fps_|ball_query|grid_|sa_mlp|sa_v2|fp_mlp|rows_to|prefix' -s 40 -c 40 -f -o gpurun_out/${R}_kernels \
    python tools/run_ops.py > gpurun_out/${R}_kernels_run.log 2>&1
ncu -i gpurun_out/${R}_kernels.ncu-rep --page raw --csv > /tmp/kernels_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/kernels_raw.csv > gpurun_out/${R}_kernels_ncu.json
;; esac
case " $STAGES " in *" train "*)
ncu --set full --import-source on --clock-control none \
    -k regex:'conv1x1|wgrad|bn_relu|bn_stats|bn_finalize|group_concat|group_points' -s 150 -c 110 -f -o gpurun_out/${R}_train_kernels \
    python tools/run_train_once.py > gpurun_out/${R}_train_kernels_run.log 2>&1
ncu -i gpurun_out/${R}_train_kernels.ncu-rep --page raw --csv > /tmp/train_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/train_raw.csv > gpurun_out/${R}_train_kernels_ncu.json
rm -f gpurun_out/${R}_train_kernels.ncu-rep      # large; the summaries are what is kept
;; esac
ls -la gpurun_out/${R}_*
