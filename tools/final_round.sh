#!/bin/bash
# Everything profiles/ needs from one 1-GPU box:  gpurun --timeout 1500 -- 'bash tools/final_round.sh'
set -x
cd "$(dirname "$0")/.."
python bench.py --steps 100 --warmup 5 > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err
python bench.py --workload detector --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/final_bench_detector.json 2> gpurun_out/final_bench_detector.err
python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/final_bench_train.json 2> gpurun_out/final_bench_train.err
python bench.py --mode train --torch-bn --steps 5 --warmup 3 > gpurun_out/final_bench_train_torchbn.json 2> gpurun_out/final_bench_train_torchbn.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
python tools/compare_ref_ext.py > gpurun_out/final_compare_ops.json 2> gpurun_out/final_compare_ops.err
python tools/compare_ref_ext.py --pipeline > gpurun_out/final_compare_pipeline.json 2> gpurun_out/final_compare_pipeline.err
python tools/sweep.py > gpurun_out/final_sweep.json 2> gpurun_out/final_sweep.err
python tools/timeline.py > gpurun_out/final_timeline.txt 2>&1
bash tools/profile_round.sh > gpurun_out/final_profile_round.log 2>&1
ls -la gpurun_out/final_* gpurun_out/kernels_ncu.json gpurun_out/launches.csv
