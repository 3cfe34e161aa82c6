"""
Synthetic `count_snps` output + genotypes of the shapes named in BASELINE.json / SURVEY.md section 8(d).

Everything is produced with numpy on the host from a seed, as molecule-level `CompressedSNPCalls` (so the row
builder is exercised) together with a `ProbabilisticGenotypes` built the way `add_vcf` would (strength 100 split
over the called alleles, genotypes.py:149-154) and a `BarcodeHandler`.  The generator deliberately includes the
cases the reference's grouping has to get right: several molecules per (variant, barcode) group in shuffled
order, multi-allelic positions whose third allele sits far away in the betas table, calls that match no
variant (sequencing errors, 'N', off-target positions), barcodes without any call, variants without any row,
denormal products of p_base_wrong, over-allocated input arrays and molecules carrying several calls.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from .barcodes import BarcodeHandler
from .calls import CompressedSNPCalls
from .genotype_store import ProbabilisticGenotypes

BASES = np.array(list('ACGT'))


@dataclass
class SyntheticDataset:
    genotypes: ProbabilisticGenotypes
    calls: Dict[str, CompressedSNPCalls]
    barcode_handler: BarcodeHandler
    barcode_donors: np.ndarray  # int [B, 2]: donor ids (second = -1 for singlets), in compressed_cb order
    n_calls: int


def donor_names(n: int) -> List[str]:
    width = max(2, len(str(n)))
    return [f'Donor{k + 1:0{width}d}' for k in range(n)]


def make_barcodes(n: int, rng: np.random.Generator) -> List[str]:
    """n distinct 16-mers with the cellranger '-1' suffix."""
    codes = rng.choice(4 ** 12, size=n, replace=False) if n <= 4 ** 11 else np.arange(n)
    out = []
    for c in codes.tolist():
        s = []
        for _ in range(16):
            s.append('ACGT'[c & 3])
            c >>= 2
        out.append(''.join(s) + '-1')
    return out


def make_dataset(n_genotypes: int, n_snps: int, n_barcodes: int, rows_per_barcode: float, seed: int,
                 n_chromosomes: int = 3, doublet_fraction: float = 0.35, third_allele_fraction: float = 0.02,
                 empty_barcode_fraction: float = 0.01, offtarget_fraction: float = 0.02,
                 tiny_error_fraction: float = 0.001, shuffle_variants: bool = False,
                 unknown_genotype_fraction: float = 0.0, spare_capacity: int = 0,
                 depth_sigma: float = 0.5, calls_seed: Optional[int] = None) -> SyntheticDataset:
    """`seed` fixes the genotypes; `calls_seed` (default: derived from `seed`) the barcodes and calls, so that
    several lanes / GPU shards can share one set of donors."""
    rng = np.random.default_rng(seed)
    G, S, B = n_genotypes, n_snps, n_barcodes

    # ---- SNP positions and alleles ----------------------------------------------------------------------
    snp_chrom = np.sort(rng.integers(0, n_chromosomes, size=S))
    snp_pos = np.empty(S, dtype=np.int64)
    for c in range(n_chromosomes):
        sel = np.flatnonzero(snp_chrom == c)
        snp_pos[sel] = np.sort(rng.choice(max(100 * len(sel), 1000), size=len(sel), replace=False)) * 7 + 11
    ref = rng.integers(0, 4, size=S)
    alt = (ref + rng.integers(1, 4, size=S)) % 4
    has_third = rng.random(S) < third_allele_fraction
    third = np.full(S, -1, dtype=np.int64)
    for s in np.flatnonzero(has_third):  # one of the two bases that are neither ref nor alt
        free = [b for b in range(4) if b != ref[s] and b != alt[s]]
        third[s] = free[int(rng.integers(0, 2))]

    # ---- donors: dosage of the alt allele, betas as add_vcf builds them --------------------------------
    freq = rng.uniform(0.05, 0.5, size=S)
    dosage = rng.binomial(2, freq[:, None], size=(S, G)).astype(np.float32)
    n_third = int(has_third.sum())
    V = 2 * S + n_third
    betas = np.zeros((V, G), dtype=np.float32)
    betas[0:2 * S:2] = 100.0 * (2 - dosage) / 2
    betas[1:2 * S:2] = 100.0 * dosage / 2
    if n_third:
        betas[2 * S:] = rng.uniform(0.0, 2.0, size=(n_third, G)).astype(np.float32)
    if unknown_genotype_fraction > 0:  # "detected SNVs": position known, genotype unknown (zero betas)
        unknown = rng.random(S) < unknown_genotype_fraction
        betas[0:2 * S:2][unknown] = 0
        betas[1:2 * S:2][unknown] = 0

    chrom_names = [f'chr{c + 1}' for c in range(n_chromosomes)]
    variant_of = np.arange(V)
    if shuffle_variants:
        variant_of = rng.permutation(V)  # logical variant k lives in row variant_of[k]
    third_snps = np.flatnonzero(has_third)
    logical_keys = [None] * V
    bases = 'ACGT'
    for s in range(S):
        name = chrom_names[snp_chrom[s]]
        pos = int(snp_pos[s])
        logical_keys[2 * s] = (name, pos, bases[ref[s]])
        logical_keys[2 * s + 1] = (name, pos, bases[alt[s]])
    for k, s in enumerate(third_snps):
        logical_keys[2 * S + k] = (chrom_names[snp_chrom[s]], int(snp_pos[s]), bases[third[s]])
    insertion = rng.permutation(V) if shuffle_variants else np.arange(V)
    genotypes = ProbabilisticGenotypes(donor_names(G))
    genotypes.var2varid = {logical_keys[k]: int(variant_of[k]) for k in insertion}
    table = np.zeros_like(betas)
    table[variant_of] = betas
    genotypes.variant_betas = table

    # ---- barcodes ---------------------------------------------------------------------------------------------
    if calls_seed is not None:
        rng = np.random.default_rng([seed, calls_seed])
    barcode_handler = BarcodeHandler(make_barcodes(B, rng))
    is_doublet = rng.random(B) < doublet_fraction
    donor_a = rng.integers(0, G, size=B)
    donor_b = (donor_a + rng.integers(1, max(G, 2), size=B)) % G
    barcode_donors = np.stack([donor_a, np.where(is_doublet & (G > 1), donor_b, -1)], axis=1)

    depth = rng.lognormal(mean=np.log(max(rows_per_barcode, 1e-9)), sigma=depth_sigma, size=B)
    depth[rng.random(B) < empty_barcode_fraction] = 0
    n_groups = rng.poisson(depth * 1.05)
    group_cb = np.repeat(np.arange(B, dtype=np.int32), n_groups)
    n_grp = len(group_cb)
    # expression skew: a few SNPs attract many molecules
    group_snp = np.minimum((S * rng.random(n_grp) ** 2.5).astype(np.int64), S - 1)
    group_snp = rng.permutation(S)[group_snp]
    per_group = rng.geometric(0.6, size=n_grp)
    call_cb = np.repeat(group_cb, per_group)
    call_snp = np.repeat(group_snp, per_group)
    M = len(call_cb)

    # ---- molecule-level calls ---------------------------------------------------------------------------------
    pick_b = (rng.random(M) < 0.5) & (barcode_donors[call_cb, 1] >= 0)
    donor = np.where(pick_b, barcode_donors[call_cb, 1], barcode_donors[call_cb, 0])
    is_alt = rng.random(M) < dosage[call_snp, donor] / 2
    base = np.where(is_alt, alt[call_snp], ref[call_snp])
    qual = rng.choice(np.array([14, 25, 37]), size=M, p=[0.05, 0.10, 0.85])
    err = (10.0 ** (-qual / 10.0)).astype(np.float32)
    second_read = rng.random(M) < 0.3  # molecules seen by two reads carry a product of two error terms
    err = np.where(second_read, err * (10.0 ** (-rng.choice(np.array([14, 25, 37]), size=M) / 10.0)).astype(np.float32),
                   err).astype(np.float32)
    flipped = rng.random(M) < np.minimum(err, 0.04)
    base = np.where(flipped, (base + rng.integers(1, 4, size=M)) % 4, base)
    base = np.where(rng.random(M) < 0.002, 4, base)  # 'N'
    tiny = rng.random(M) < tiny_error_fraction
    err = np.where(tiny, np.float32(3e-30), err).astype(np.float32)  # products of these reach denormals / zero
    position = snp_pos[call_snp] + np.where(rng.random(M) < offtarget_fraction, 1, 0)
    chrom = snp_chrom[call_snp]

    calls: Dict[str, CompressedSNPCalls] = {}
    for c in range(n_chromosomes):
        sel = np.flatnonzero(chrom == c)
        sel = sel[rng.permutation(len(sel))]
        if len(sel) == 0 and c > 0:
            continue
        # a molecule may carry two calls: merge ~10% of neighbours that share the barcode
        order = sel[np.argsort(call_cb[sel], kind='stable')]
        same_cb = np.zeros(len(order), dtype=bool)
        same_cb[1:] = call_cb[order][1:] == call_cb[order][:-1]
        starts_molecule = ~(same_cb & (rng.random(len(order)) < 0.1))
        if len(order):
            starts_molecule[0] = True
        mol_of_call = np.cumsum(starts_molecule) - 1
        n_mol = int(mol_of_call[-1]) + 1 if len(order) else 0
        mol_perm = rng.permutation(n_mol)  # molecule ids in arbitrary order
        mol_of_call = mol_perm[mol_of_call] if n_mol else mol_of_call
        mol_cb = np.zeros(n_mol, dtype=np.int32)
        mol_cb[mol_of_call] = call_cb[order]
        shuffle = rng.permutation(len(order))  # calls in arbitrary order as well
        order, mol_of_call = order[shuffle], mol_of_call[shuffle]
        calls[chrom_names[c]] = CompressedSNPCalls.from_arrays(
            compressed_cb=mol_cb, compressed_ub=rng.integers(0, 2 ** 31 - 1, size=n_mol),
            p_group_misaligned=rng.uniform(1e-4, 0.1, size=n_mol).astype(np.float32),
            molecule_index=mol_of_call, snp_position=position[order], base_index=base[order],
            p_base_wrong=err[order], spare_capacity=spare_capacity)
    return SyntheticDataset(genotypes=genotypes, calls=calls, barcode_handler=barcode_handler,
                            barcode_donors=barcode_donors, n_calls=M)


# Named workloads; `scale` < 1 shrinks barcodes and SNPs for parity-sized runs.
CONFIGS = {
    'example_like': dict(n_genotypes=4, n_snps=1212, n_barcodes=1000, rows_per_barcode=1180, doublet_fraction=0.25),
    'pbmc_32': dict(n_genotypes=32, n_snps=325_000, n_barcodes=10_000, rows_per_barcode=1600),
    'em_32_3m': dict(n_genotypes=32, n_snps=1_500_000, n_barcodes=10_000, rows_per_barcode=3000,
                     unknown_genotype_fraction=0.78),
    'biobank_200': dict(n_genotypes=200, n_snps=2_500_000, n_barcodes=100_000, rows_per_barcode=5000),
    'lane_64': dict(n_genotypes=64, n_snps=650_000, n_barcodes=12_500, rows_per_barcode=2000),
}


def make_config(name: str, scale: float = 1.0, seed: Optional[int] = None, **overrides) -> SyntheticDataset:
    cfg = dict(CONFIGS[name])
    if scale != 1.0:
        cfg['n_snps'] = max(16, int(cfg['n_snps'] * scale))
        cfg['n_barcodes'] = max(8, int(cfg['n_barcodes'] * scale))
    cfg.update(overrides)
    if seed is None:
        seed = 20260000 + list(CONFIGS).index(name)
    return make_dataset(seed=seed, **cfg)
