#!/bin/bash
# Round 2, final one-GPU evidence: GPU test suite, default bench, ncu launch list of the bench command, ncu --set full
# of the dominant kernel, the small-width flavour comparison.
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -q -m gpu) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log
(time timeout 1200 python bench.py) > gpurun_out/r02_bench.log 2>&1
grep -c '^{' gpurun_out/r02_bench.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:estep_pairs_strip -c 1 -f -o gpurun_out/r02_prof_strip \
    python scripts/profile_estep.py 1.0 1 0.35 pbmc_32 auto > gpurun_out/r02_ncu_strip.log 2>&1
timeout 300 python scripts/bench_flavours_small.py > gpurun_out/r02_flavours_small.log 2>&1
cat gpurun_out/r02_flavours_small.log
