"""Rebuilds the input objects of a golden fixture (tests/golden/*.npz, written by make_golden.py)."""
from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace

import numpy as np

from demuxalot_b200 import BarcodeHandler, CompressedSNPCalls, ProbabilisticGenotypes

GOLDEN_DIR = Path(__file__).resolve().parent / 'golden'
# synthetic cases (make_golden.py) + config #1: the reference's bundled example, first 48 barcodes (make_example_fixture.py)
CASES = ['g4_dp25', 'g7_dp0_prior', 'g12_dp35', 'g33_dp35_lowdepth', 'example_data_48bc']


def load_case(name: str) -> SimpleNamespace:
    fx = dict(np.load(GOLDEN_DIR / f'{name}.npz', allow_pickle=False))
    genotypes = ProbabilisticGenotypes([str(g) for g in fx['genotype_names']], default_prior=float(fx['default_prior']))
    genotypes.var2varid = {
        (str(c), int(p), str(b)): int(v)
        for c, p, b, v in zip(fx['var_chrom'], fx['var_pos'], fx['var_base'], fx['var_id'])
    }
    genotypes.variant_betas = fx['raw_betas'].copy()
    calls = {}
    for chrom in fx['chromosomes']:
        chrom = str(chrom)
        c = CompressedSNPCalls.__new__(CompressedSNPCalls)
        c.molecules, c.snp_calls = fx[f'mol__{chrom}'], fx[f'calls__{chrom}']
        c.n_molecules, c.n_snp_calls = (int(x) for x in fx[f'n__{chrom}'])
        calls[chrom] = c
    barcode_handler = BarcodeHandler([str(b) for b in fx['barcodes']])
    assert barcode_handler.ordered_barcodes == [str(b) for b in fx['barcodes']]
    return SimpleNamespace(
        name=name, fx=fx, genotypes=genotypes, calls=calls, barcode_handler=barcode_handler,
        doublet_prior=float(fx['doublet_prior']), p_genotype_clip=float(fx['p_genotype_clip']),
        n_iterations=int(fx['n_iterations']), prior_logits=fx.get('prior_logits'))


def bits(a: np.ndarray) -> np.ndarray:
    """Bit pattern view for exact float comparison (distinguishes -0/+0 and denormals, equates NaNs)."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    if a.dtype == np.float64:
        return a.view(np.uint64)
    return a
