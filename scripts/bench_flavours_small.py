"""Row-stream-bound E-step paths (at most 8 genotypes, or singlet columns only): the reference's per-term roundings
(DMX_ESTEP_EXACT, what DMX_ESTEP_AUTO picks there) against the product arithmetic (DMX_ESTEP_FAST), device-generated rows.
    python scripts/bench_flavours_small.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from demuxalot_b200 import Demultiplexer as D
from demuxalot_b200.synthetic_device import make_device_dataset

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
for G, dps in ((4, (0.35, 0.0)), (8, (0.35, 0.0)), (32, (0.0,)), (200, (0.0,))):
    B, rows = (10_000, 2000) if G <= 32 else (4_000, 3000)
    ds = make_device_dataset(n_genotypes=G, n_snps=300_000, n_barcodes=B, rows_per_barcode=rows, seed=79)
    part = ds.device_calls(np.arange(B), 'cuda')
    pack = D._pack_device(None, ds.genotypes, B, add_data_prior=False, device_parts=[part], keep_calls=False)
    del part
    table = D._probs_table(pack, None, 0.01)
    for dp in dps:
        C = G * (G + 1) // 2 if dp else G
        bytes_ = pack.n_rows * (8 + 4 * G) + 4 * B * C
        line = f'G={G:3d} dp={dp:4.2f} C={C:5d} R={pack.n_rows}: '
        for flavour in ('fast', 'exact'):
            D.estep_flavour = flavour
            pack.__dict__.pop('_estep_plans', None)
            buffers, times = {}, []
            for i in range(8):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                D._e_step(pack, table, dp, want_logits=True, want_post=False, buffers=buffers)
                b.record()
                torch.cuda.synchronize()
                if i > 1:
                    times.append(a.elapsed_time(b))
            best = min(times)
            line += f' {flavour} {best:7.3f} ms = {bytes_ / best / 1e6:6.0f} GB/s algorithmic ({100 * bytes_ / best / 1e6 / 6543:3.0f} % of HBM peak);'
        print(line, flush=True)
    D.estep_flavour = 'auto'
    del pack, table, ds
    torch.cuda.empty_cache()
