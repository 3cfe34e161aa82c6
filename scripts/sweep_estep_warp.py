"""Times the warp-per-item pair E-step (csrc/estep_pairs_warp.cu) against the CTA-per-barcode kernel on the bench
workload:  python scripts/sweep_estep_warp.py [scale] [workload] [key=value ...]
Sweeps the segment length (Demultiplexer.estep_segment_rows) and the DMX_WARP_VARIANT experiments."""
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
print(f'G={G} C={C} B={pack.n_barcodes} R={pack.n_rows} updates={pack.n_rows * C:.3e}', flush=True)


def run(seg_rows, env, reps=7, want_post=False):
    for k in ('DMX_WARP_VARIANT', 'DMX_FLUSH_ROWS', 'DMX_PAIRS_WARP'):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    Demultiplexer.estep_segment_rows = seg_rows
    pack.__dict__.pop('_estep_plans', None)
    buffers = {}
    times = []
    for i in range(reps + 2):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        logits, _, _ = Demultiplexer._e_step(pack, table, 0.35, want_logits=True, want_post=want_post, buffers=buffers)
        b.record()
        torch.cuda.synchronize()
        if i > 1:
            times.append(a.elapsed_time(b))
    plan = Demultiplexer._estep_plan(pack, 0.35)
    return min(times), sum(times) / len(times), logits.clone(), (plan[2] if plan else 0)


base = None
cases = [(0, {})] + [(4096, dict(DMX_WARP_VARIANT=v)) for v in (0, 1, 2)] + \
        [(2048, dict(DMX_WARP_VARIANT=0)), (1024, dict(DMX_WARP_VARIANT=0)), (4096, dict(DMX_FLUSH_ROWS=8))]
for seg_rows, env in cases:
    try:
        best, mean, logits, n_items = run(seg_rows, env)
    except Exception as exc:  # noqa: BLE001
        print(seg_rows, env, 'FAILED', exc, flush=True)
        continue
    if base is None:
        base = logits
    d = (logits.double() - base.double()).abs()
    rel = (d / base.double().abs().clamp_min(1e-30)).max().item()
    print(f'seg_rows={seg_rows:5d} items={n_items:6d} {env}  best {best:.3f} ms  mean {mean:.3f} ms  '
          f'{pack.n_rows * C / best / 1e9:.2f} T upd/s  {pack.n_rows * C / (best * 1e-3) / 148 / 1.965e9:.1f} upd/clk/SM  '
          f'max rel dlogit vs CTA kernel {rel:.2e}', flush=True)
