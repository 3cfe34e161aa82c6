"""Wall-clock breakdown of Demultiplexer.predict_posteriors on the bench workload: python scripts/profile_e2e.py [scale]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import pandas as pd
import torch

from bench import pin
from demuxalot_b200 import Demultiplexer
from demuxalot_b200.demultiplexer import _to_device, option_names
from demuxalot_b200.synthetic import make_config

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ds = make_config('pbmc_32', scale=scale)
pinned = all([pin(c.snp_calls) and pin(c.molecules) for c in ds.calls.values()] + [pin(ds.genotypes.variant_betas)])
print('pinned', pinned)
dev = torch.device('cuda', 0)


def timed(label, fn, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f'{label:45s} {1e3 * best:9.3f} ms')
    return out


timed('hot_path_index (cached after first)', lambda: ds.genotypes.hot_path_index())
raw = np.asarray(ds.genotypes.get_betas())
timed('host: raw.min() >= 0', lambda: raw.min() >= 0)
total = 0
for chrom, c in ds.calls.items():
    total += c.snp_calls[:c.n_snp_calls].nbytes + c.molecules[:c.n_molecules].nbytes
    timed(f'H2D snp_calls {chrom} ({c.snp_calls[:c.n_snp_calls].nbytes / 1e6:.0f} MB)', lambda: _to_device(c.snp_calls[:c.n_snp_calls], dev))
    timed(f'H2D molecules {chrom} ({c.molecules[:c.n_molecules].nbytes / 1e6:.0f} MB)', lambda: _to_device(c.molecules[:c.n_molecules], dev))
timed(f'H2D betas ({raw.nbytes / 1e6:.0f} MB)', lambda: _to_device(raw, dev))
print(f'total call bytes {total / 1e6:.0f} MB')
B = ds.barcode_handler.n_barcodes
pack = timed('_pack_device (H2D + match + build + prior)', lambda: Demultiplexer._pack_device(ds.calls, ds.genotypes, B, False))
table = timed('_probs_table', lambda: Demultiplexer._probs_table(pack, None, 0.01))
logits, post, _ = timed('_e_step (pairs + softmax)', lambda: Demultiplexer._e_step(pack, table, 0.35))
ln, pn = timed('D2H logits + post (.cpu().numpy())', lambda: (logits.cpu().numpy(), post.cpu().numpy()))
pinned_out = torch.empty((2,) + tuple(logits.shape), dtype=torch.float32).pin_memory()


def d2h_pinned():
    pinned_out[0].copy_(logits, non_blocking=True)
    pinned_out[1].copy_(post, non_blocking=True)
    torch.cuda.synchronize()
    return pinned_out.numpy()


timed('D2H into pinned buffer', d2h_pinned)
names = option_names(ds.genotypes.genotype_names, 0.35)
index = list(ds.barcode_handler.ordered_barcodes)
timed('2 x pd.DataFrame', lambda: (pd.DataFrame(ln, index=index, columns=names), pd.DataFrame(pn, index=index, columns=names)))
timed('predict_posteriors end to end', lambda: Demultiplexer.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=0.35))

# inside _pack_device: kernels only (inputs already resident)
import ctypes as C
from demuxalot_b200 import _native
lib = _native.load()
idx = ds.genotypes.hot_path_index()
gkeys, gvids = _to_device(idx['keys_sorted'], dev), _to_device(idx['vids_sorted'], dev)
dcalls = {k: (_to_device(c.snp_calls[:c.n_snp_calls], dev), _to_device(c.molecules[:c.n_molecules], dev), c) for k, c in ds.calls.items()}
n_calls = sum(c.n_snp_calls for c in ds.calls.values())
cv = torch.empty(n_calls, dtype=torch.int32, device=dev); cc = torch.empty_like(cv); ce = torch.empty(n_calls, dtype=torch.float32, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def unpack_all():
    done = 0
    for k, (a, m, c) in dcalls.items():
        lib.dmx_unpack_match_calls(a.data_ptr(), c.n_snp_calls, m.data_ptr(), c.n_molecules, 12, idx['chrom2id'][k], gkeys.data_ptr(),
                                   gvids.data_ptr(), len(idx['keys_sorted']), cv[done:].data_ptr(), cc[done:].data_ptr(), ce[done:].data_ptr(), stream)
        done += c.n_snp_calls


timed('kernels: unpack + match (all chromosomes)', unpack_all)
V = ds.genotypes.n_variants
ws_bytes = lib.dmx_build_rows_workspace_bytes(n_calls, V, B)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
bufs = [torch.empty(n_calls, dtype=torch.int32, device=dev) for _ in range(6)] + [torch.empty(n_calls, dtype=torch.float32, device=dev) for _ in range(2)]
vo = torch.empty(V + 1, dtype=torch.int64, device=dev); bo = torch.empty(B + 1, dtype=torch.int64, device=dev); nm = torch.empty(V, dtype=torch.int64, device=dev)
hr, hm = C.c_int64(0), C.c_int64(0)
timed('kernels: build rows (2 sorts, scan, products)', lambda: lib.dmx_build_rows(
    cv.data_ptr(), cc.data_ptr(), ce.data_ptr(), n_calls, V, B, 0, B, ws.data_ptr(), ws_bytes, bufs[0].data_ptr(), bufs[1].data_ptr(),
    bufs[6].data_ptr(), bufs[2].data_ptr(), vo.data_ptr(), bufs[3].data_ptr(), bufs[7].data_ptr(), bufs[4].data_ptr(), bo.data_ptr(),
    nm.data_ptr(), C.byref(hr), C.byref(hm), stream))
print('rows', hr.value, 'workspace MB', ws_bytes / 1e6)
