"""Times the pair E-step under the DMX_* tuning overrides (csrc/estep_pairs.cu): python scripts/sweep_estep.py [scale] [workload]"""
import itertools
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from demuxalot_b200 import Demultiplexer
from demuxalot_b200.synthetic import make_config

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
workload = sys.argv[2] if len(sys.argv) > 2 else 'pbmc_32'
overrides = {}
for kv in sys.argv[3:]:
    k, v = kv.split('=')
    overrides[k] = float(v) if '.' in v else int(v)
ds = make_config(workload, scale=scale, **overrides)
pack = Demultiplexer._pack_device(ds.calls, ds.genotypes, ds.barcode_handler.n_barcodes, add_data_prior=False)
table = Demultiplexer._probs_table(pack, None, 0.01)
G = pack.n_genotypes
C = G * (G + 1) // 2
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
print(f'G={G} C={C} B={pack.n_barcodes} R={pack.n_rows} updates={pack.n_rows * C:.3e}')


def run(env, reps=5):
    for k in ('DMX_RG', 'DMX_FLUSHES', 'DMX_FLUSH_ROWS', 'DMX_MAX_THREADS', 'DMX_MIN_BLOCKS', 'DMX_INT_WIDEN', 'DMX_TJ'):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    buffers = {}
    times = []
    for i in range(reps + 1):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        logits, _, _ = Demultiplexer._e_step(pack, table, 0.35, want_logits=True, want_post=False, buffers=buffers)
        b.record()
        torch.cuda.synchronize()
        if i:
            times.append(a.elapsed_time(b))
    return min(times), sum(times) / len(times), logits.clone()


base = None
grid = dict(DMX_TJ=[4, 8], DMX_MAX_THREADS=[128, 256], DMX_FLUSH_ROWS=[16], DMX_FLUSHES=[1, 2])
for combo in itertools.product(*grid.values()):
    env = dict(zip(grid.keys(), combo), DMX_VERBOSE=0)
    try:
        best, mean, logits = run(env)
    except Exception as exc:  # noqa: BLE001
        print(env, 'FAILED', exc)
        continue
    if base is None:
        base = logits
    diff = (logits.double() - base.double()).abs().max().item()
    print(f'{env}  best {best:.3f} ms  mean {mean:.3f} ms  {pack.n_rows * C / best / 1e9:.2f} T upd/s  '
          f'max|dlogit| vs first {diff:.2e}')
