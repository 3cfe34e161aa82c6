"""
Device-side synthetic `count_snps` output for workloads too large to draw on the host (bench / test tooling).

BASELINE.json configs[3] ("low-depth biobank": 200 donors, 100 k barcodes, 5 M variants, 500 M read rows) holds about
800 M molecule-level calls; numpy needs minutes and tens of GB for that.  Here every random decision is a pure integer
function of (seed, barcode, group, molecule): `csrc/synth.cu` evaluates it on the GPU for any
set of barcodes -- so every GPU count sees the same data set -- and `host_calls()` below evaluates the same function
with numpy uint64 arithmetic for a small barcode subset, which is what the CPU oracle is given.

The donors / SNPs / barcodes themselves (small: O(S * G) and O(B)) are drawn with numpy from the seed on every rank.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Optional

import numpy as np

from .barcodes import BarcodeHandler
from .calls import CompressedSNPCalls
from .genotype_store import ProbabilisticGenotypes
from .synthetic import donor_names

CSRC = Path(__file__).resolve().parent / 'csrc'
SYNTH_LIB_PATH = CSRC / 'libdemux_synth.so'
_U64 = np.uint64
_QUALITIES = np.array([14, 25, 37])
CHROMOSOME = 'chr1'


def build_synth(force: bool = False) -> Path:
    from .build import NVCC_FLAGS, find_nvcc
    src = CSRC / 'synth.cu'
    if not force and SYNTH_LIB_PATH.exists() and SYNTH_LIB_PATH.stat().st_mtime >= src.stat().st_mtime:
        return SYNTH_LIB_PATH
    flags = [f for f in NVCC_FLAGS if f not in ('-Xptxas', '-v')]
    cmd = [find_nvcc(), *flags, '-shared', str(src), '-o', str(SYNTH_LIB_PATH)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed on synth.cu:\n{" ".join(cmd)}\n{res.stdout}')
    return SYNTH_LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(build_synth()))
        p, i64, i32, u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64
        lib.dmxs_count.argtypes = [u64, i64, i32, p, p, i64, i64, p, p]
        lib.dmxs_emit.argtypes = [u64, i64, i32, p, p, i64, i64, p, p, p, p, p, p, p, p, p, p, p, p, p, p, p]
        lib.dmxs_pack.argtypes = [p, i64, p, p, p, p, p, p, p]
        for fn in (lib.dmxs_count, lib.dmxs_emit, lib.dmxs_pack):
            fn.restype = C.c_int
        _lib = lib
    return _lib


# ------------------------------------------------------------------------------------------------ integer hashing
def _mix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(_U64, copy=True)
    x ^= x >> _U64(30)
    x *= _U64(0xbf58476d1ce4e5b9)
    x ^= x >> _U64(27)
    x *= _U64(0x94d049bb133111eb)
    x ^= x >> _U64(31)
    return x


def _draw(seed: int, barcode: np.ndarray, group: np.ndarray, what) -> np.ndarray:
    with np.errstate(over='ignore'):
        inner = _mix64(_U64(seed) + barcode.astype(_U64) * _U64(0x9e3779b97f4a7c15))
        return _mix64(inner + group.astype(_U64) * _U64(0xd1b54a32d192ed03) + np.asarray(what, dtype=_U64))


def _molecules_in_group(g2: np.ndarray) -> np.ndarray:
    m = (g2 >> _U64(32)) & _U64(0xffff)
    return np.where(m < 39322, 1, np.where(m < 55050, 2, np.where(m < 61342, 3, 4))).astype(np.int64)


# ------------------------------------------------------------------------------------------------ the data set
@dataclass
class DeviceSyntheticDataset:
    """Donors, SNPs and barcodes of a device-generated workload (the calls are generated on demand)."""
    seed: int
    genotypes: ProbabilisticGenotypes
    barcode_handler: BarcodeHandler
    barcode_donors: np.ndarray      # int32 [B, 2], second = -1 for singlets
    groups_per_barcode: np.ndarray  # int64 [B] (variant, barcode) groups drawn per barcode (before collisions)
    dosage: np.ndarray              # int8 [S, G]
    ref_base: np.ndarray            # uint8 [S]
    alt_base: np.ndarray            # uint8 [S]
    snp_position: np.ndarray        # int32 [S]
    err_table: np.ndarray           # float32 [12]
    flip_threshold: np.ndarray      # int32 [12]

    @property
    def n_snps(self) -> int:
        return len(self.ref_base)

    # -- host mirror of csrc/synth.cu -------------------------------------------------------------------------
    def host_calls(self, barcodes: np.ndarray, relabel: bool = True) -> Dict[str, CompressedSNPCalls]:
        """The calls of the given (global) barcode ids as the input stage would deliver them; with `relabel` the
        compressed_cb are 0..len(barcodes)-1 in the given order (a self-contained slice for the oracle)."""
        barcodes = np.asarray(barcodes, dtype=np.int64)
        n_groups = self.groups_per_barcode[barcodes]
        b = np.repeat(barcodes, n_groups)
        local = np.repeat(np.arange(len(barcodes)), n_groups)
        grp = np.arange(len(b)) - np.repeat(np.cumsum(n_groups) - n_groups, n_groups)
        g1, g2 = _draw(self.seed, b, grp, 0), _draw(self.seed, b, grp, 1)
        S = _U64(self.n_snps)
        t = ((g1 & _U64(0xffffffff)) * (g1 >> _U64(32))) >> _U64(32)
        t = (t * (g2 & _U64(0xffffffff))) >> _U64(32)
        rank = (t * S) >> _U64(32)
        snp = ((rank * _U64(2654435761) + _U64(12345)) % S).astype(np.int64)
        m = _molecules_in_group(g2)
        # one entry per molecule
        j = np.arange(int(m.sum())) - np.repeat(np.cumsum(m) - m, m)
        b, grp, snp, local = (np.repeat(x, m) for x in (b, grp, snp, local))
        c = _draw(self.seed, b, grp, 2 + j)
        c2 = _draw(self.seed, b, grp, 64 + j)
        f0, f1 = c & _U64(0xffff), (c >> _U64(16)) & _U64(0xffff)
        f2, f3 = (c >> _U64(32)) & _U64(0xffff), (c >> _U64(48)) & _U64(0xffff)
        da, db = self.barcode_donors[b, 0], self.barcode_donors[b, 1]
        donor = np.where((db >= 0) & (f1 < 32768), db, da)
        dose = self.dosage[snp, donor].astype(np.uint64)
        base = np.where(f0 < dose * _U64(32768), self.alt_base[snp], self.ref_base[snp]).astype(np.int64)
        q1 = np.where(f2 < 3277, 0, np.where(f2 < 9830, 1, 2))
        err_index = np.where(f3 < 19661, 3 + 3 * q1 + (f3 % _U64(3)).astype(np.int64), q1)
        flip = (c2 & _U64(0xfffff)).astype(np.int64) < self.flip_threshold[err_index]
        shift = 1 + (((c2 >> _U64(20)) & _U64(0xff)) % _U64(3)).astype(np.int64)
        base = np.where(flip, (base + shift) & 3, base)
        base = np.where(((c2 >> _U64(28)) & _U64(0xffff)) < 131, 4, base)
        off_target = (((c2 >> _U64(44)) & _U64(0xffff)) < 1311).astype(np.int64)
        key = (_mix64(c ^ _U64(0x9e3779b97f4a7c15)) >> _U64(1)).astype(np.int64)
        order = np.argsort(key, kind='stable')
        n = len(order)
        cb = (local if relabel else b)[order]
        return {CHROMOSOME: CompressedSNPCalls.from_arrays(
            compressed_cb=cb.astype(np.int32), compressed_ub=np.zeros(n, dtype=np.int32),
            p_group_misaligned=np.full(n, 0.01, dtype=np.float32), molecule_index=np.arange(n, dtype=np.int32),
            snp_position=(self.snp_position[snp] + off_target)[order], base_index=base[order].astype(np.uint8),
            p_base_wrong=self.err_table[err_index][order])}

    # -- device generation -------------------------------------------------------------------------------------------
    def device_tables(self, device) -> dict:
        import torch
        cache = self.__dict__.setdefault('_device_tables', {})
        key = torch.device(device).index
        if key not in cache:
            up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
            cache[key] = dict(donor_a=up(self.barcode_donors[:, 0].astype(np.int32)),
                              donor_b=up(self.barcode_donors[:, 1].astype(np.int32)), dosage=up(self.dosage),
                              ref=up(self.ref_base), alt=up(self.alt_base), pos=up(self.snp_position),
                              err=up(self.err_table), flip=up(self.flip_threshold))
        return cache[key]

    def device_calls(self, barcodes: np.ndarray, device) -> dict:
        """Calls of the given (global) barcode ids generated on `device`, in the form the uploads of the row builder
        leave them in: packed 13-byte snp_calls records (uint8 [n, 13]) + the compressed_cb column of the molecules
        (int32 [n]; one molecule per call), compressed_cb = global barcode id."""
        import torch
        lib = _load()
        dev = torch.device(device)
        t = self.device_tables(dev)
        barcodes = np.asarray(barcodes, dtype=np.int32)
        prefix = np.concatenate([[0], np.cumsum(self.groups_per_barcode[barcodes])]).astype(np.int64)
        n_groups = int(prefix[-1])
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            d_ids = torch.from_numpy(barcodes).to(dev)
            d_prefix = torch.from_numpy(prefix).to(dev)
            molecules = torch.empty(max(n_groups, 1), dtype=torch.int32, device=dev)
            G = self.genotypes.n_genotypes
            rc = lib.dmxs_count(self.seed, self.n_snps, G, d_ids.data_ptr(), d_prefix.data_ptr(), len(barcodes),
                                n_groups, molecules.data_ptr(), stream)
            assert rc == 0, 'dmxs_count failed'
            offsets = torch.cumsum(molecules[:n_groups], dim=0, dtype=torch.int64)
            n_calls = int(offsets[-1]) if n_groups else 0
            offsets -= molecules[:n_groups]
            del molecules
            pos = torch.empty(max(n_calls, 1), dtype=torch.int32, device=dev)
            base = torch.empty(max(n_calls, 1), dtype=torch.uint8, device=dev)
            e = torch.empty(max(n_calls, 1), dtype=torch.float32, device=dev)
            cb = torch.empty(max(n_calls, 1), dtype=torch.int32, device=dev)
            key = torch.empty(max(n_calls, 1), dtype=torch.int64, device=dev)
            rc = lib.dmxs_emit(self.seed, self.n_snps, G, d_ids.data_ptr(), d_prefix.data_ptr(), len(barcodes), n_groups,
                               t['donor_a'].data_ptr(), t['donor_b'].data_ptr(), t['dosage'].data_ptr(),
                               t['ref'].data_ptr(), t['alt'].data_ptr(), t['pos'].data_ptr(), t['err'].data_ptr(),
                               t['flip'].data_ptr(), offsets.data_ptr(), pos.data_ptr(), base.data_ptr(), e.data_ptr(),
                               cb.data_ptr(), key.data_ptr(), stream)
            assert rc == 0, 'dmxs_emit failed'
            del offsets
            perm = torch.argsort(key[:n_calls], stable=True)
            del key
            records = torch.empty((max(n_calls, 1), 13), dtype=torch.uint8, device=dev)
            molecule_cb = torch.empty(max(n_calls, 1), dtype=torch.int32, device=dev)
            rc = lib.dmxs_pack(perm.data_ptr(), n_calls, pos.data_ptr(), base.data_ptr(), e.data_ptr(), cb.data_ptr(),
                               records.data_ptr(), molecule_cb.data_ptr(), stream)
            assert rc == 0, 'dmxs_pack failed'
            torch.cuda.current_stream().synchronize()
        return dict(chromosome=CHROMOSOME, records=records[:n_calls], molecule_cb=molecule_cb[:n_calls], n_calls=n_calls)


def make_device_dataset(n_genotypes: int, n_snps: int, n_barcodes: int, rows_per_barcode: float, seed: int,
                        doublet_fraction: float = 0.35, depth_sigma: float = 0.5,
                        unknown_genotype_fraction: float = 0.0, calls_seed: Optional[int] = None) -> DeviceSyntheticDataset:
    """`seed` fixes donors and SNPs; `calls_seed` (default: derived from `seed`) the barcodes and their calls, so that
    several lanes can share one set of donors."""
    rng = np.random.default_rng(seed)
    G, S, B = n_genotypes, n_snps, n_barcodes
    snp_position = (11 + 37 * np.arange(S, dtype=np.int64)).astype(np.int32)
    assert 11 + 37 * S < 2 ** 31
    ref = rng.integers(0, 4, size=S).astype(np.uint8)
    alt = ((ref + rng.integers(1, 4, size=S)) % 4).astype(np.uint8)
    freq = rng.uniform(0.05, 0.5, size=S)
    dosage = np.empty((S, G), dtype=np.int8)
    betas = np.zeros((2 * S, G), dtype=np.float32)
    step = max(1, (1 << 24) // max(G, 1))
    for lo in range(0, S, step):  # chunks keep the int64 temporaries of rng.binomial small
        hi = min(S, lo + step)
        d = rng.binomial(2, freq[lo:hi, None], size=(hi - lo, G)).astype(np.int8)
        dosage[lo:hi] = d
        betas[2 * lo:2 * hi:2] = 50.0 * (2 - d)  # add_vcf: strength 100 split over the two called alleles
        betas[2 * lo + 1:2 * hi:2] = 50.0 * d
    if unknown_genotype_fraction > 0:  # "detected SNVs": position known, genotype unknown
        unknown = np.flatnonzero(rng.random(S) < unknown_genotype_fraction)
        betas[2 * unknown] = 0
        betas[2 * unknown + 1] = 0
    genotypes = ProbabilisticGenotypes(donor_names(G))
    bases = 'ACGT'
    positions = snp_position.tolist()
    var2varid = {}
    for s, (pos, r, a) in enumerate(zip(positions, ref.tolist(), alt.tolist())):
        var2varid[(CHROMOSOME, pos, bases[r])] = 2 * s
        var2varid[(CHROMOSOME, pos, bases[a])] = 2 * s + 1
    genotypes.var2varid = var2varid
    genotypes.variant_betas = betas

    hash_seed = int(seed)
    if calls_seed is not None:
        rng = np.random.default_rng([seed, calls_seed])
        hash_seed = int(seed) * 1_000_003 + int(calls_seed) + 1
    width = len(str(B))
    barcode_handler = BarcodeHandler([f'BC{k:0{width}d}-1' for k in range(B)])  # already sorted: id k <-> name k
    is_doublet = rng.random(B) < doublet_fraction
    donor_a = rng.integers(0, G, size=B)
    donor_b = (donor_a + rng.integers(1, max(G, 2), size=B)) % G
    barcode_donors = np.stack([donor_a, np.where(is_doublet & (G > 1), donor_b, -1)], axis=1).astype(np.int32)
    depth = rng.lognormal(mean=np.log(max(rows_per_barcode, 1e-9)), sigma=depth_sigma, size=B)
    depth[rng.random(B) < 0.01] = 0  # empty droplets
    groups = rng.poisson(depth * 1.02).astype(np.int64)

    single = (10.0 ** (-_QUALITIES / 10.0)).astype(np.float32)
    err_table = np.concatenate([single, (single[:, None] * single[None, :]).astype(np.float32).ravel()]).astype(np.float32)
    flip_threshold = np.floor(np.minimum(err_table.astype(np.float64), 0.04) * 2 ** 20).astype(np.int32)
    return DeviceSyntheticDataset(seed=hash_seed, genotypes=genotypes, barcode_handler=barcode_handler,
                                  barcode_donors=barcode_donors, groups_per_barcode=groups, dosage=dosage,
                                  ref_base=ref, alt_base=alt, snp_position=snp_position, err_table=err_table,
                                  flip_threshold=flip_threshold)


DEVICE_CONFIGS = {
    # BASELINE.json configs[3]
    # rows_per_barcode counts (SNP, barcode) groups: ~1.2 (variant, barcode) rows each -> ~500 M read rows
    'biobank_200': dict(n_genotypes=200, n_snps=2_500_000, n_barcodes=100_000, rows_per_barcode=4050),
}


def make_device_config(name: str, scale: float = 1.0, seed: Optional[int] = None, **overrides) -> DeviceSyntheticDataset:
    cfg = dict(DEVICE_CONFIGS[name])
    if scale != 1.0:
        cfg['n_snps'] = max(64, int(cfg['n_snps'] * scale))
        cfg['n_barcodes'] = max(8, int(cfg['n_barcodes'] * scale))
    cfg.update(overrides)
    return make_device_dataset(seed=20260003 if seed is None else seed, **cfg)
