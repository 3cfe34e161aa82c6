#!/bin/bash
# Round 2, first GPU pass: parity tests, the pinned microbenchmark, the default bench (headline + biobank_200 + em_32_3m).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02a_gpu.txt 2>&1
(time timeout 1500 python -m pytest tests -q -m gpu --durations=15) > gpurun_out/r02a_pytest.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench_packed_tile scripts/microbench_packed_tile.cu && \
    timeout 300 scripts/microbench_packed_tile > gpurun_out/r02a_microbench_packed_tile.log 2>&1
(time timeout 1200 python bench.py) > gpurun_out/r02a_bench.log 2>&1
tail -25 gpurun_out/r02a_pytest.log; tail -c 6000 gpurun_out/r02a_bench.log
