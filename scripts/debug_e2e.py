"""Per-call wall times of predict_posteriors under the bench's conditions (resident pack, flush buffer)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from bench import pin
from demuxalot_b200 import Demultiplexer
from demuxalot_b200.synthetic import make_config

ds = make_config('pbmc_32', scale=1.0, calls_seed=0)
pinned = all([pin(c.snp_calls) and pin(c.molecules) for c in ds.calls.values()] + [pin(ds.genotypes.variant_betas)])
ds.genotypes.hot_path_index()
B = ds.barcode_handler.n_barcodes


def loop(tag, n=8):
    times = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a, b = Demultiplexer.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=0.35)
        torch.cuda.synchronize()
        times.append(1e3 * (time.perf_counter() - t0))
    print(f'{tag:50s}', ' '.join(f'{t:6.1f}' for t in times), flush=True)


import gc
gc.collect()
gc.freeze()
for threads in (0, 4, 8, 16):
    Demultiplexer.host_gather_threads = threads
    loop(f'host_gather_threads={threads}', n=10)
Demultiplexer.host_gather_threads = 8
Demultiplexer.pipelined_upload = False
loop('host_gather_threads=8, serial upload', n=6)
