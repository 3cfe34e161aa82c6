#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python scripts/sweep_estep_strip.py 32 64 > gpurun_out/r02d_sweep_strip.log 2>&1
cat gpurun_out/r02d_sweep_strip.log
DMX_STRIP_UNROLL=${1:-16} timeout 600 ncu --set full --clock-control none --import-source on -k regex:estep_pairs_strip -c 6 -f -o gpurun_out/r02d_prof_strip \
    python scripts/sweep_estep_strip.py 32 > gpurun_out/r02d_ncu_strip.log 2>&1
tail -2 gpurun_out/r02d_ncu_strip.log
