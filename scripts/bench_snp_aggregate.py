"""Times the `aggregate_on_snps = True` E-step (dmx_build_snp_groups once, then dmx_snp_logits + dmx_softmax_rows_f64)
beside the default E-step on the same pack, CUDA events on the launching stream:
python scripts/bench_snp_aggregate.py [scale] [workload]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from demuxalot_b200 import Demultiplexer
from demuxalot_b200.demultiplexer import n_options
from demuxalot_b200.synthetic import make_config

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
workload = sys.argv[2] if len(sys.argv) > 2 else 'pbmc_32'
dp = 0.35
ds = make_config(workload, scale=scale)
pack = Demultiplexer._pack_device(ds.calls, ds.genotypes, ds.barcode_handler.n_barcodes, add_data_prior=True)
table = Demultiplexer._probs_table(pack, None, 0.01)
torch.cuda.synchronize()
t0 = time.perf_counter()
*_, n_matched, n_groups = Demultiplexer._snp_groups(pack)
torch.cuda.synchronize()
t_groups = time.perf_counter() - t0
n_cols = n_options(pack.n_genotypes, dp)


def timed(fn, reps=3):
    fn()
    times = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return sorted(times)[len(times) // 2]


buffers, buffers_default = {}, {}
ms = timed(lambda: Demultiplexer._e_step_aggregated(pack, table, dp, want_singlets=True, buffers=buffers))
ms_default = timed(lambda: Demultiplexer._e_step(pack, table, dp, want_singlets=True, buffers=buffers_default))
print(f'{workload} x {scale}: B={pack.n_barcodes} G={pack.n_genotypes} C={n_cols} matched calls={n_matched} '
      f'groups={n_groups} rows={pack.n_rows}')
print(f'group builder (once per pack, wall clock incl. allocation) {1e3 * t_groups:8.2f} ms')
print(f'aggregate_on_snps E-step {ms:9.3f} ms  {n_matched * n_cols / ms / 1e6:8.2f} G call x column logs/s  '
      f'{n_groups * n_cols / ms / 1e6:8.2f} G group x column/s')
print(f'default E-step           {ms_default:9.3f} ms  {pack.n_rows * n_cols / ms_default / 1e6:8.2f} G row x column updates/s')
