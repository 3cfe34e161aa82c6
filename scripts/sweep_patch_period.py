"""Patch kernel (G = 200): flush period 16 against the opt-in period 32 (operands staged times 4), device-generated rows.
    python scripts/sweep_patch_period.py [G]"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from demuxalot_b200 import Demultiplexer as D
from demuxalot_b200.synthetic_device import make_device_dataset

G = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B = 3000
ds = make_device_dataset(n_genotypes=G, n_snps=200_000, n_barcodes=B, rows_per_barcode=3000, seed=78)
part = ds.device_calls(np.arange(B), 'cuda')
pack = D._pack_device(None, ds.genotypes, B, add_data_prior=False, device_parts=[part], keep_calls=False)
del part
table = D._probs_table(pack, None, 0.01)
C = G * (G + 1) // 2
sms = torch.cuda.get_device_properties(0).multi_processor_count
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
print(f'G={G} C={C} B={B} R={pack.n_rows} updates={pack.n_rows * C:.3e}', flush=True)
base = None
for label, env in (('period 16', {}), ('period 32, 168 regs', dict(DMX_PATCH_PERIOD='32')),
                   ('period 32, 184 regs', dict(DMX_PATCH_PERIOD='32', DMX_PATCH_REGS='184'))):
    for k in ('DMX_PATCH_PERIOD', 'DMX_PATCH_REGS'):
        os.environ.pop(k, None)
    os.environ.update(env)
    buffers, times = {}, []
    for i in range(6):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        logits, _, _ = D._e_step(pack, table, 0.35, want_logits=True, want_post=False, buffers=buffers)
        b.record()
        torch.cuda.synchronize()
        if i > 1:
            times.append(a.elapsed_time(b))
    logits = logits.clone()
    base = logits if base is None else base
    rel = ((logits.double() - base.double()).abs() / base.double().abs().clamp_min(1e-30)).max().item()
    print(f'  {label:22s} best {min(times):8.3f} ms  {pack.n_rows * C / (min(times) * 1e-3) / sms / 1.965e9:5.1f} upd/clk/SM  '
          f'max rel dlogit vs period 16 {rel:.2e}', flush=True)
