#!/bin/bash
# strip kernel: parity tests, sweep against the previous kernels, ncu capture
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "warp_pair or shapes or predict or learn") > gpurun_out/r02c_pytest.log 2>&1
tail -5 gpurun_out/r02c_pytest.log
timeout 900 python scripts/sweep_estep_strip.py 32 64 27 61 > gpurun_out/r02c_sweep_strip.log 2>&1
cat gpurun_out/r02c_sweep_strip.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:estep_pairs_strip -c 2 -f -o gpurun_out/r02c_prof_strip \
    python scripts/sweep_estep_strip.py 32 > gpurun_out/r02c_ncu_strip.log 2>&1
tail -3 gpurun_out/r02c_ncu_strip.log
