"""Strip kernel (csrc/estep_pairs_strip.cu) against the kernels it replaces, on device-generated rows:
    python scripts/sweep_estep_strip.py [G ...]
For every width: strip kernel with flush periods 32 / 16, the previous kernel (DMX_PAIRS_STRIP=0: warp kernel at G <= 32,
CTA kernel at 57..64), time per E-step, updates/clk/SM and the largest relative logit difference between the two."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from demuxalot_b200 import Demultiplexer as D
from demuxalot_b200.synthetic_device import make_device_dataset

widths = [int(a) for a in sys.argv[1:]] or [32, 64, 27, 61]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
sms = torch.cuda.get_device_properties(0).multi_processor_count
for G in widths:
    rows = 2000 if G <= 32 else 1500
    B = 10_000 if G <= 32 else 6_000
    ds = make_device_dataset(n_genotypes=G, n_snps=300_000, n_barcodes=B, rows_per_barcode=rows, seed=77)
    part = ds.device_calls(np.arange(B), 'cuda')
    pack = D._pack_device(None, ds.genotypes, B, add_data_prior=False, device_parts=[part], keep_calls=False)
    del part
    table = D._probs_table(pack, None, 0.01)
    C = G * (G + 1) // 2
    print(f'G={G} C={C} B={B} R={pack.n_rows} updates={pack.n_rows * C:.3e}', flush=True)
    base = None
    cases = [('previous kernel', dict(DMX_PAIRS_STRIP='0')), ('strip, flush period 16', dict(DMX_STRIP_PERIOD='16'))]
    cases += [(f'strip, period 32, regs {r}, unroll {u}', dict(DMX_STRIP_REGS=str(r), DMX_STRIP_UNROLL=str(u)))
              for u in (16, 8) for r in (168, 184, 200)]
    for label, env in cases:
        for k in ('DMX_PAIRS_STRIP', 'DMX_STRIP_PERIOD', 'DMX_STRIP_UNROLL', 'DMX_STRIP_REGS'):
            os.environ.pop(k, None)
        os.environ.update(env)
        pack.__dict__.pop('_estep_plans', None)
        buffers, times = {}, []
        for i in range(9):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            logits, _, _ = D._e_step(pack, table, 0.35, want_logits=True, want_post=False, buffers=buffers)
            b.record()
            torch.cuda.synchronize()
            if i > 1:
                times.append(a.elapsed_time(b))
        logits = logits.clone()
        if base is None:
            base = logits
        rel = ((logits.double() - base.double()).abs() / base.double().abs().clamp_min(1e-30)).max().item()
        best = min(times)
        print(f'  {label:40s} best {best:.3f} ms  mean {sum(times) / len(times):.3f} ms  '
              f'{pack.n_rows * C / (best * 1e-3) / sms / 1.965e9:.1f} upd/clk/SM  max rel dlogit vs previous {rel:.2e}',
              flush=True)
    for k in ('DMX_PAIRS_STRIP', 'DMX_STRIP_PERIOD', 'DMX_STRIP_UNROLL', 'DMX_STRIP_REGS'):
        os.environ.pop(k, None)
    del pack, table, logits, base, buffers
    torch.cuda.empty_cache()
