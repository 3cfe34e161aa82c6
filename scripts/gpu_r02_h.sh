#!/bin/bash
# one GPU: patch-kernel period sweep, full GPU test suite, default bench (strip kernel now serves the headline width)
set -u
mkdir -p gpurun_out
timeout 600 python scripts/sweep_patch_period.py 200 > gpurun_out/r02h_sweep_patch_period.log 2>&1
cat gpurun_out/r02h_sweep_patch_period.log
(time timeout 1500 python -m pytest tests -q -m gpu) > gpurun_out/r02h_pytest.log 2>&1
tail -4 gpurun_out/r02h_pytest.log
(time timeout 1200 python bench.py) > gpurun_out/r02h_bench.log 2>&1
tail -c 7000 gpurun_out/r02h_bench.log
