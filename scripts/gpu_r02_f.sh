#!/bin/bash
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "warp_pair or shapes") > gpurun_out/r02f_pytest.log 2>&1
tail -3 gpurun_out/r02f_pytest.log
timeout 900 python scripts/sweep_estep_strip.py 32 64 > gpurun_out/r02f_sweep_strip.log 2>&1
cat gpurun_out/r02f_sweep_strip.log
