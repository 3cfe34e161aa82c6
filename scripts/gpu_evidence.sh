#!/bin/bash
# One-GPU evidence run: tests, bench (both arms), ncu launch list of the bench command, ncu --set full captures,
# e2e breakdown, compute-sanitizer on smoke().  Outputs go to gpurun_out/ (scripts/collect_profiles.py files them).
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.log 2>&1
python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:estep_pairs_warp -c 1 -f -o gpurun_out/prof_estep_full \
    python scripts/profile_estep.py 1.0 1 > gpurun_out/ncu_estep.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:mstep_tiers|probs_table|softmax_rows' -c 3 -f -o gpurun_out/prof_aux \
    python scripts/profile_estep.py 1.0 1 > gpurun_out/ncu_aux.log 2>&1
ncu --set full --clock-control none -k regex:estep_singlets -c 1 -f -o gpurun_out/prof_singlets \
    python scripts/profile_estep.py 1.0 1 0.0 > gpurun_out/ncu_singlets.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:estep_pairs_patch -c 1 -f -o gpurun_out/prof_patch \
    python scripts/sweep_estep_patch.py 200 > gpurun_out/ncu_patch.log 2>&1
python scripts/run_config4_shard.py > gpurun_out/config4_shard.log 2>&1
python scripts/profile_e2e.py > gpurun_out/profile_e2e.log 2>&1
python scripts/bench_mstep.py > gpurun_out/bench_mstep.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench_packed_tile scripts/microbench_packed_tile.cu && \
    scripts/microbench_packed_tile > gpurun_out/microbench_packed_tile.log 2>&1
(python scripts/sweep_table.py 32 656584; python scripts/sweep_table.py 200 1000000) > gpurun_out/sweep_table.log 2>&1
python scripts/bench_snp_aggregate.py 0.25 > gpurun_out/bench_snp_aggregate.log 2>&1
# aggregate_on_snps E-step: tens of ms per launch and ~40 replays -- a tenth of the workload and a generous limit
timeout 300 ncu --set full --clock-control none --import-source on -k regex:snp_logits -c 1 -f -o gpurun_out/prof_snp_logits \
    python scripts/bench_snp_aggregate.py 0.1 > gpurun_out/ncu_snp_logits.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke 2>&1 | tail -15 > gpurun_out/sanitizer.log
tail -3 gpurun_out/pytest_gpu.log; tail -c 1500 gpurun_out/bench.log; echo; tail -c 600 gpurun_out/bench_reference.log; echo; tail -4 gpurun_out/sanitizer.log
