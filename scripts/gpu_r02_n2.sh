#!/bin/bash
# Round 2, two GPUs: the 2-GPU parity tests (sharded pack + dmx_mstep_allreduce vs one GPU and vs the oracle), then the
# default bench under torchrun (headline weak scaling + biobank_200 strong scaling + lanes_64 + multi_gpu_parity).
set -u
mkdir -p gpurun_out
N=${1:-2}
(time timeout 600 python -m pytest tests/test_distributed.py -q -m gpu) > gpurun_out/r02_pytest_dist_n$N.log 2>&1
(time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5) > gpurun_out/r02_bench_n$N.log 2>&1
tail -5 gpurun_out/r02_pytest_dist_n$N.log; tail -c 7000 gpurun_out/r02_bench_n$N.log
