"""
One GPU's share of BASELINE config #4 (low-depth biobank: 200 donors, 20 100 columns, 5 M variants, 100 k barcodes,
500 M rows over 8 GPUs -> 12.5 k barcodes / ~62 M rows per GPU): 10 EM iterations of learn_genotypes, stage timings.

    python scripts/run_config4_shard.py [n_barcodes] [n_snps] [rows_per_barcode] [n_iterations]
"""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from demuxalot_b200 import Demultiplexer
from demuxalot_b200.synthetic import make_dataset

n_barcodes = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500
n_snps = int(sys.argv[2]) if len(sys.argv) > 2 else 2_500_000
rows_per_barcode = float(sys.argv[3]) if len(sys.argv) > 3 else 4000
n_iterations = int(sys.argv[4]) if len(sys.argv) > 4 else 10
G, dp = 200, 0.35

t0 = time.perf_counter()
ds = make_dataset(n_genotypes=G, n_snps=n_snps, n_barcodes=n_barcodes, rows_per_barcode=rows_per_barcode,
                  seed=20260003, tiny_error_fraction=0.0)
t_gen = time.perf_counter() - t0
print(f'generated {ds.n_calls} calls, {ds.genotypes.n_variants} variants in {t_gen:.1f} s', flush=True)

torch.cuda.synchronize()
t0 = time.perf_counter()
ds.genotypes.hot_path_index()
t_index = time.perf_counter() - t0
t0 = time.perf_counter()
pack = Demultiplexer._pack_device(ds.calls, ds.genotypes, n_barcodes, add_data_prior=True)
torch.cuda.synchronize()
t_pack = time.perf_counter() - t0
C = G * (G + 1) // 2
print(f'index {t_index:.1f} s, pack {t_pack:.2f} s: rows {pack.n_rows}, updates/iteration {pack.n_rows * C:.3e}', flush=True)

events = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
events[0].record()
post, addition = Demultiplexer._em_iterations(pack, n_iterations, 0.01, dp, None)
events[1].record()
torch.cuda.synchronize()
em_ms = events[0].elapsed_time(events[1])
assert bool(torch.isfinite(post).all()) and bool(torch.isfinite(addition).all())
row_sums = post.sum(dim=1)
assert bool(((row_sums - 1).abs() < 1e-3).all())
called = post[:, :G].argmax(dim=1).cpu().numpy()
conf = post[:, :G].max(dim=1).values.cpu().numpy() > 0.9
singlet = ds.barcode_donors[:, 1] < 0
acc = float((called[conf & singlet] == ds.barcode_donors[conf & singlet, 0]).mean())
out = dict(
    workload='config #4 shard (1 of 8 GPUs)', donors=G, columns=C, variants=pack.n_variants, barcodes=n_barcodes,
    molecule_calls=pack.n_calls, rows=pack.n_rows, n_iterations=n_iterations,
    host_generation_s=t_gen, host_index_s=t_index, pack_s=t_pack, em_ms=em_ms,
    em_iterations_per_s=n_iterations / (em_ms / 1e3),
    estep_updates_per_s=pack.n_rows * C * n_iterations / (em_ms / 1e3),
    confident_singlet_fraction=float((conf & singlet).sum() / max(singlet.sum(), 1)), confident_singlet_accuracy=acc,
    peak_device_memory_gb=torch.cuda.max_memory_allocated() / 1e9,
)
print(json.dumps(out))
