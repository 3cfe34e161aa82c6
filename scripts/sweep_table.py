"""Sweep of the probability-table kernel's knobs (DMX_TABLE_MAXV / MINB / CAP) through the C ABI, CUDA-event timed
with an L2 flush between launches: python scripts/sweep_table.py [G] [V]"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from demuxalot_b200 import _native

G = int(sys.argv[1]) if len(sys.argv) > 1 else 32
V = int(sys.argv[2]) if len(sys.argv) > 2 else 656_584
lib = _native.load()
dev = torch.device('cuda:0')
gen = torch.Generator(device=dev).manual_seed(1)
betas = torch.rand(V, G, device=dev, generator=gen) * 100 + 1
addition = torch.rand(V, G, device=dev, generator=gen) * 10
offsets = torch.arange(0, V + 1, 2, device=dev, dtype=torch.int32)  # bi-allelic SNPs
variants = torch.arange(V, device=dev, dtype=torch.int32)
S = offsets.numel() - 1
table = torch.empty(V, G, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def launch():
    _native.check(lib.dmx_probs_from_betas(betas.data_ptr(), G, addition.data_ptr(), G, V, G, offsets.data_ptr(),
                                           variants.data_ptr(), S, 0.01, 0.99, table.data_ptr(), G, stream), 'table')


def timed(reps=20):
    for _ in range(3):
        launch()
    times = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        launch()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    times.sort()
    return times[len(times) // 2]


gb = (3 * V * G * 4 + V * 4) / 1e9
reference = None
for maxv, minb, cap in [(4, 1, 32), (4, 4, 32), (2, 4, 32), (2, 5, 32), (4, 1, 0), (4, 4, 0), (2, 4, 0), (2, 5, 0),
                        (4, 4, 16), (2, 4, 16), (2, 5, 16), (2, 4, 8), (2, 5, 8), (4, 4, 8), (2, 4, 64), (4, 4, 64)]:
    os.environ.update(DMX_TABLE_MAXV=str(maxv), DMX_TABLE_MINB=str(minb), DMX_TABLE_CAP=str(cap))
    ms = timed()
    if reference is None:
        reference = table.clone()
    same = bool(torch.equal(reference, table))
    print(f'G={G} V={V} maxv={maxv} minb={minb} cap={cap:3d}  {ms * 1e3:7.1f} us  {gb / ms:7.1f} TB/s  bit-identical={same}',
          flush=True)
