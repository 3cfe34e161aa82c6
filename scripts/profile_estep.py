"""Runs the resident hot path a few times (for ncu / timing experiments): python scripts/profile_estep.py [scale] [reps] [dp]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from demuxalot_b200 import Demultiplexer
from demuxalot_b200.synthetic import make_config

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.35
workload = sys.argv[4] if len(sys.argv) > 4 else 'pbmc_32'
flavour = sys.argv[5] if len(sys.argv) > 5 else 'fast'
Demultiplexer.estep_flavour = flavour
ds = make_config(workload, scale=scale)
pack = Demultiplexer._pack_device(ds.calls, ds.genotypes, ds.barcode_handler.n_barcodes, add_data_prior=True)
table = Demultiplexer._probs_table(pack, None, 0.01)
buffers = {}
addition = torch.zeros_like(pack.betas)
for _ in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Demultiplexer._probs_table(pack, addition, 0.01, out=table)
    _, _, singlets = Demultiplexer._e_step(pack, table, dp, want_logits=True, want_post=True, want_singlets=True,
                                           buffers=buffers)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    addition = Demultiplexer._m_step(pack, singlets)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'rows={pack.n_rows} table+estep {1e3 * (t1 - t0):.3f} ms, mstep {1e3 * (t2 - t1):.3f} ms')
