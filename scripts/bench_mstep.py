"""Times the M-step schedules on the bench workload: python scripts/bench_mstep.py [scale] [workload]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from demuxalot_b200 import Demultiplexer
from demuxalot_b200.synthetic import make_config

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
workload = sys.argv[2] if len(sys.argv) > 2 else 'pbmc_32'
overrides = {kv.split('=')[0]: int(kv.split('=')[1]) for kv in sys.argv[3:]}
ds = make_config(workload, scale=scale, **overrides)
pack = Demultiplexer._pack_device(ds.calls, ds.genotypes, ds.barcode_handler.n_barcodes, add_data_prior=True)
table = Demultiplexer._probs_table(pack, None, 0.01)
_, _, singlets = Demultiplexer._e_step(pack, table, 0.35, want_logits=False, want_post=False, want_singlets=True)
G, V, R = pack.n_genotypes, pack.n_variants, pack.n_rows
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
out = torch.empty((V, G), dtype=torch.float32, device='cuda')
algorithmic = R * (8 + 4 * G) + V * G * 4
print(f'G={G} V={V} R={R} algorithmic bytes {algorithmic / 1e9:.2f} GB', flush=True)
results = {}
for planned in (False, True, False, True):
    Demultiplexer.planned_mstep = planned
    times = []
    for i in range(9):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = Demultiplexer._m_step(pack, singlets, out=out)
        b.record()
        torch.cuda.synchronize()
        if i > 1:
            times.append(a.elapsed_time(b))
    results[planned] = res.clone()
    plan = Demultiplexer._mstep_plan(pack) if planned else None
    print(f'planned={planned}  best {min(times):.3f} ms  mean {sum(times) / len(times):.3f} ms  '
          f'{algorithmic / min(times) / 1e6:.0f} GB/s algorithmic' + (f'  tiers medium={plan[1]} heavy={plan[2]} items={plan[3]}' if plan else ''),
          flush=True)
print('identical:', torch.equal(results[False], results[True]))
