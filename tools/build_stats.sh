#!/bin/bash
# Developer build of the library with the SA phase trace compiled in (tools/sa_trace.py):
#   bash tools/build_stats.sh   ->  tools/bin/libbqa_stats.so
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr \
     -DBQA_SA_TRACE -shared -o tools/bin/libbqa_stats.so bridgeqa_b200/csrc/*.cu
