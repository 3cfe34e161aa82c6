"""Patch kernel (csrc/estep_pairs_warp.cu, many genotypes) against the CTA-per-barcode kernel:
python scripts/sweep_estep_patch.py [G ...]"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from demuxalot_b200 import Demultiplexer
from demuxalot_b200.synthetic import make_dataset

widths = [int(a) for a in sys.argv[1:]] or [200, 104]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
for G in widths:
    C = G * (G + 1) // 2
    n_barcodes = max(300, int(float(os.environ.get('SWEEP_UPDATES', '2.0e11')) / C / 3000))  # thousands of work items
    ds = make_dataset(n_genotypes=G, n_snps=100_000, n_barcodes=n_barcodes, rows_per_barcode=3000, seed=20260003,
                      tiny_error_fraction=0.0)
    pack = Demultiplexer._pack_device(ds.calls, ds.genotypes, n_barcodes, add_data_prior=False)
    table = Demultiplexer._probs_table(pack, None, 0.01)
    print(f'G={G} C={C} B={n_barcodes} R={pack.n_rows} V={pack.n_variants} updates={pack.n_rows * C:.3e}', flush=True)
    base = None
    for label, env, seg in (('CTA kernel', dict(DMX_PAIRS_PATCH=0), 4096), ('patch kernel', dict(DMX_PAIRS_PATCH=1, DMX_PAIRS_PATCH_MIN_EFF=os.environ.get('SWEEP_MIN_EFF', '80')), 4096),
                            ('patch kernel seg 2048', dict(DMX_PAIRS_PATCH=1), 2048)):
        os.environ.update({k: str(v) for k, v in env.items()})
        Demultiplexer.estep_segment_rows = seg
        pack.__dict__.pop('_estep_plans', None)
        buffers, times = {}, []
        for i in range(5):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            logits, _, _ = Demultiplexer._e_step(pack, table, 0.35, want_logits=True, want_post=False, buffers=buffers)
            b.record()
            torch.cuda.synchronize()
            if i:
                times.append(a.elapsed_time(b))
        if base is None:
            base = logits.clone()
        rel = ((logits.double() - base.double()).abs() / base.double().abs().clamp_min(1e-30)).max().item()
        best = min(times)
        print(f'  {label:24s} best {best:8.3f} ms  {pack.n_rows * C / best / 1e9:7.2f} T upd/s  '
              f'{pack.n_rows * C / (best * 1e-3) / 148 / 1.965e9:5.1f} upd/clk/SM  max rel dlogit vs CTA {rel:.2e}', flush=True)
    del pack, table, ds
