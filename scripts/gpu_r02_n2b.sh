#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
(time timeout 900 python -m pytest tests -q -m gpu) > gpurun_out/r02i_pytest_n$N.log 2>&1
tail -3 gpurun_out/r02i_pytest_n$N.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    scripts/sweep_mstep_allreduce.py) > gpurun_out/r02_sweep_mstep_allreduce_n$N.log 2>&1
grep -v "^\*\|OMP_NUM" gpurun_out/r02_sweep_mstep_allreduce_n$N.log | tail -40
(time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 20 --warmup 5) > gpurun_out/r02i_bench_n$N.log 2>&1
tail -c 6500 gpurun_out/r02i_bench_n$N.log
