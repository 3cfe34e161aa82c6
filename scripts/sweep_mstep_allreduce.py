"""EM iteration of the headline workload (pbmc_32, one shard per rank) with the cross-GPU sum of dmx_mstep_allreduce:
tiles x wire format, against the same iteration without the sum.  Run under torchrun:
    python -m torch.distributed.run --nproc-per-node N scripts/sweep_mstep_allreduce.py [scale]
Also times torch's symmetric-memory multimem all-reduce on the same buffer when the platform offers it (NVLS)."""
import json
import os
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

from demuxalot_b200 import Demultiplexer as D
from demuxalot_b200.synthetic import make_config

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=dev)
ds = make_config('pbmc_32', scale=scale, calls_seed=rank)
pack = D._pack_device(ds.calls, ds.genotypes, ds.barcode_handler.n_barcodes, add_data_prior=True, keep_calls=False)
V, G = pack.n_variants, pack.n_genotypes
table = D._probs_table(pack, None, 0.01)
buffers = {}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, reps=15):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        out.append((a, b))
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([statistics.mean(x.elapsed_time(y) for x, y in out)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def iteration(mbuf, state):
    cur, nxt = mbuf['tables'][state['k'] & 1], mbuf['tables'][(state['k'] + 1) & 1]
    D._probs_table(pack, cur[:V], 0.01, out=table)
    _, _, singlets = D._e_step(pack, table, 0.35, want_logits=False, want_post=False, want_singlets=True, buffers=buffers)
    D._m_step(pack, singlets, out=nxt, buffers=mbuf)
    state['k'] += 1


result = {'world': world, 'V': V, 'G': G, 'rows_per_rank': pack.n_rows}
D.process_group = None
mbuf, state = D._mstep_buffers(pack), {'k': 0}
result['em_local_ms'] = timed(lambda: iteration(mbuf, state))
D.process_group = dist.group.WORLD
D.mstep_exchange = 'peer'
mbuf, state = D._mstep_buffers(pack), {'k': 0}
result['exchange_peer_available'] = 'peer' in mbuf
if 'peer' in mbuf:
    result['em_ms/peer (dmx_peer_sum_f32)'] = timed(lambda: iteration(mbuf, state))
    assert mbuf.get('last_exchange') == 'peer/float32', mbuf.get('last_exchange')
D.mstep_exchange = 'nccl'
for wire in ('float64', 'float32'):
    for tiles in (1, 2, 4, 8):
        D.mstep_allreduce_dtype, D.mstep_allreduce_tiles = wire, tiles
        mbuf, state = D._mstep_buffers(pack), {'k': 0}
        ms = timed(lambda: iteration(mbuf, state))
        result[f'em_ms/{mbuf.get("last_exchange")}'] = ms  # keyed by what RAN, not by what was asked for
        assert mbuf.get('last_exchange') == f'nccl/{wire}/tiles={tiles}', (mbuf.get('last_exchange'), wire, tiles)
# plain NCCL collectives on the same buffer, alone
buf32 = torch.zeros((V, G), dtype=torch.float32, device=dev)
buf64 = torch.zeros((V, G), dtype=torch.float64, device=dev)
result['nccl_allreduce_f32_ms'] = timed(lambda: dist.all_reduce(buf32))
result['nccl_allreduce_f64_ms'] = timed(lambda: dist.all_reduce(buf64))
try:
    import torch.distributed._symmetric_memory as symm_mem
    n = (V * G + 1023) // 1024 * 1024
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    t.zero_()
    for name in ('multimem_all_reduce_', 'two_shot_all_reduce_', 'one_shot_all_reduce'):
        try:
            op = getattr(torch.ops.symm_mem, name)
            result[f'symm_mem/{name}_f32_ms'] = timed(lambda: op(t, 'sum', dist.group.WORLD.group_name))
        except Exception as exc:  # noqa: BLE001
            result[f'symm_mem/{name}'] = f'failed: {exc!r}'[:300]
except Exception as exc:  # noqa: BLE001
    result['symm_mem'] = f'unavailable: {exc!r}'[:300]
if rank == 0:
    print(json.dumps(result, indent=1))
from demuxalot_b200.distributed import release_native_comms
release_native_comms()
dist.destroy_process_group()
