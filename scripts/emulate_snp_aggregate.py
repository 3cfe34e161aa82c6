"""CPU emulation (numpy, scalar loops) of the index logic and pass structure of csrc/snp_aggregate.cu, checked against
the oracle on a golden case with unmatched calls mixed in (the kernel now adds the groups of a barcode per warp and
combines 8 partial rows; this emulation keeps the sequential order: same values up to float64 regrouping).  A development aid for a container without a GPU:
python scripts/emulate_snp_aggregate.py"""
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'tests')]
import numpy as np

import oracle
from golden_io import load_case

case = load_case('g12_dp35')
O = oracle.OracleDemultiplexer
v2s, betas, mol, rows = O.pack_calls(case.calls, case.genotypes, False)
table = oracle.probs_from_betas(v2s, betas, case.p_genotype_clip)
B = case.barcode_handler.n_barcodes
n_snps = int(v2s.max()) + 1
rng = np.random.default_rng(0)
M, extra = len(mol['variant_id']), 500
n = M + extra
mask = np.zeros(n, bool)
mask[np.sort(rng.choice(n, M, replace=False))] = True  # matched calls keep their relative order
cv, cc, ce = np.full(n, -1, np.int32), rng.integers(0, B, n).astype(np.int32), np.zeros(n, np.float32)
cv[mask], cc[mask], ce[mask] = mol['variant_id'], mol['compressed_cb'], mol['p_base_wrong']

# dmx_build_snp_groups
sentinel = B * n_snps
keys = np.where(cv >= 0, cc.astype(np.int64) * n_snps + v2s[np.maximum(cv, 0)], sentinel)
order = np.argsort(keys, kind='stable')
ks = keys[order]
flags = np.ones(n, np.int32)
flags[1:] = ks[1:] != ks[:-1]
incl = np.cumsum(flags)
gvar, ge, go = np.zeros(n, np.int32), np.zeros(n, np.float32), np.full(n + 1, -7, np.int64)
for k in range(n):
    if ks[k] != sentinel:
        gvar[k], ge[k] = cv[order[k]], ce[order[k]]
    if k == 0 or ks[k] != ks[k - 1]:
        go[incl[k] - 1] = k
bgo = np.zeros(B + 1, np.int64)
for b in range(B + 1):
    lo = int(np.searchsorted(ks, b * n_snps, side='left'))
    group = incl[lo] - 1 if lo < n else incl[n - 1]
    bgo[b] = group
    if b == B:
        go[group] = lo
        n_matched, n_groups = lo, group
_, counts, group_barcode = oracle.snp_groups(mol['compressed_cb'], mol['snp_id'])
assert n_matched == M and n_groups == len(counts), (n_matched, M, n_groups, len(counts))
assert np.array_equal(np.diff(go[:n_groups + 1]), counts)
assert np.array_equal(bgo, np.searchsorted(group_barcode, np.arange(B + 1)))
print('groups ok:', n_groups)

# snp_pairs_kernel
dp, G = case.doublet_prior, table.shape[1]
pairs = oracle.demux_oracle.option_pairs(G, dp)
C = len(pairs)


def pair_of(c):
    i = j = c
    if c >= G:
        d, i = c - G, 0
        while d >= G - 1 - i:
            d -= G - 1 - i
            i += 1
        j = i + 1 + d
    return i, j


assert all(pair_of(c) == tuple(pairs[c]) for c in range(C))

# snp_logits_kernel
f32 = np.float32
log_bad = math.log(0.01 / C)


def logaddexp(x, y):
    if x == y:
        return x + math.log(2)
    d = x - y
    return x + math.log1p(math.exp(-d)) if d > 0 else y + math.log1p(math.exp(d))


logits = np.zeros((B, C))
for b in range(B):
    for q in range(bgo[b], bgo[b + 1]):
        lo, hi = go[q], go[q + 1]
        denom = math.sqrt(hi - lo)
        xs = np.zeros(C, f32)
        for c in range(C):
            i, j = pairs[c]
            s = 0.0
            for k in range(lo, hi):
                row = table[gvar[k]]
                p = row[i] if i == j else f32(f32(row[i] + row[j]) * f32(0.5))
                s += float(np.log(f32(p + ge[k])))
            xs[c] = f32(float(f32(s)) / denom)
        m = xs.max()
        l = np.log(f32(np.exp(xs - m).sum()))
        z_max = logaddexp(float(f32(0) - l), log_bad)
        ts = np.array([logaddexp(float(f32(f32(x - m) - l)), log_bad) - z_max for x in xs])
        logits[b] += ts - math.log(np.exp(ts).sum())
want = oracle.snp_aggregated_logits(mol['variant_id'], mol['snp_id'], mol['compressed_cb'], mol['p_base_wrong'],
                                    table, dp, B)
print('logits: max |diff|', np.abs(logits - want).max(), 'at max |logit|', np.abs(want).max())
assert np.abs(logits - want).max() < 1e-6
