"""
Multi-GPU EM iteration: how to cut and type the variant x genotype all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/sweep_allreduce.py [--workload pbmc_32] [--out profiles/allreduce_sweep.json]

Every rank holds its own barcode shard of the workload (weak scaling, as in bench.py).  Timed with CUDA events on
the launching stream, max over ranks: the M-step alone, the bare NCCL all-reduce of the [V, G] buffer in float64
and float32, and the whole EM iteration for each (tiles, dtype) setting of Demultiplexer.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='pbmc_32')
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--out', default='')
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    assert world > 1, 'run under torchrun with at least two ranks'
    dist.init_process_group('nccl', device_id=dev)
    from demuxalot_b200 import Demultiplexer
    from demuxalot_b200.synthetic import make_config

    ds = make_config(args.workload, scale=args.scale, calls_seed=rank)
    B = ds.barcode_handler.n_barcodes
    pack = Demultiplexer._pack_device(ds.calls, ds.genotypes, B, add_data_prior=False)
    table = Demultiplexer._probs_table(pack, None, 0.01)
    buffers: dict = {}
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def timed(fn, iters=args.iters, warm=3):
        for _ in range(warm):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        total = 0.0
        for _ in range(iters):
            flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            total += a.elapsed_time(b)
        t = torch.tensor([total / iters], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    results = {'world': world, 'workload': args.workload, 'V': pack.n_variants, 'G': pack.n_genotypes,
               'rows_per_rank': pack.n_rows}
    _, _, singlets = Demultiplexer._e_step(pack, table, 0.35, want_logits=False, want_post=False, want_singlets=True,
                                           buffers=buffers)
    spare = torch.empty_like(pack.betas)
    spare64 = torch.empty(pack.betas.shape, dtype=torch.float64, device=dev)
    Demultiplexer.process_group = None
    results['mstep_ms'] = timed(lambda: Demultiplexer._m_step(pack, singlets, out=spare))
    for name, buf in (('float64', spare64), ('float32', spare)):
        buf.zero_()
        results[f'allreduce_{name}_ms'] = timed(lambda: dist.all_reduce(buf))
        halves = [buf[:pack.n_variants // 4], buf[pack.n_variants // 4:pack.n_variants // 2]]
        results[f'allreduce_{name}_quarter_ms'] = timed(lambda: dist.all_reduce(halves[0]))

    state = {'addition': torch.zeros_like(pack.betas), 'spare': spare}

    def em_iteration():
        Demultiplexer._probs_table(pack, state['addition'], 0.01, out=table)
        _, _, s = Demultiplexer._e_step(pack, table, 0.35, want_logits=False, want_post=False, want_singlets=True,
                                        buffers=buffers)
        new = Demultiplexer._m_step(pack, s, out=state['spare'], out64=spare64)
        state['spare'], state['addition'] = state['addition'], new

    results['em_local_ms'] = timed(em_iteration)
    Demultiplexer.process_group = dist.group.WORLD
    results['em_ms'] = {}
    for dtype in ('float64', 'float32'):
        for tiles in (1, 2, 4, 8):
            Demultiplexer.mstep_allreduce_dtype, Demultiplexer.mstep_allreduce_tiles = dtype, tiles
            state['addition'].zero_()
            results['em_ms'][f'{dtype}/tiles={tiles}'] = timed(em_iteration)
    Demultiplexer.process_group = None
    if rank == 0:
        text = json.dumps(results, indent=1)
        print(text)
        if args.out:
            Path(args.out).parent.mkdir(parents=True, exist_ok=True)
            Path(args.out).write_text(text + '\n')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
