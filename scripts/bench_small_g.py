"""Pair and singlet E-step at small donor counts (the reference's own example has 4): time against the HBM roofline.
python scripts/bench_small_g.py [G ...]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from demuxalot_b200 import Demultiplexer
from demuxalot_b200.synthetic import make_config

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
for G in [int(a) for a in sys.argv[1:]] or [4, 8, 16, 48, 64]:
    ds = make_config('pbmc_32', n_genotypes=G)
    pack = Demultiplexer._pack_device(ds.calls, ds.genotypes, ds.barcode_handler.n_barcodes, add_data_prior=False)
    table = Demultiplexer._probs_table(pack, None, 0.01)
    for dp in (0.35, 0.0):
        C = G * (G + 1) // 2 if dp else G
        buffers, times = {}, []
        for i in range(6):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            Demultiplexer._e_step(pack, table, dp, want_logits=True, want_post=False, buffers=buffers)
            b.record()
            torch.cuda.synchronize()
            if i:
                times.append(a.elapsed_time(b))
        best = min(times)
        bytes_ = pack.n_rows * (8 + 4 * G) + 4 * pack.n_barcodes * C
        print(f'G={G:3d} dp={dp:4.2f} C={C:5d} R={pack.n_rows}  {best:7.3f} ms  {pack.n_rows * C / best / 1e9:7.2f} T upd/s  '
              f'{pack.n_rows * C / (best * 1e-3) / 148 / 1.965e9:5.1f} upd/clk/SM  {bytes_ / best / 1e6:7.0f} GB/s algorithmic '
              f'({100 * bytes_ / best / 1e6 / 6543:.0f} % of HBM peak)', flush=True)
    del pack, table, ds
