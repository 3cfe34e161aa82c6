"""
Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, demuxalot v0.4.3, pysam
stubbed -- see tests/reference_loader.py) on small seeded synthetic inputs.  Run in the build container:

    python tests/golden/make_golden.py

Each fixture stores the flattened inputs (so tests can rebuild the exact objects without the generator) and the
reference's outputs for: pack_calls (rows, ids, p_base_wrong bits, regularised betas with and without the data
prior), _compute_probs_from_betas, predict_posteriors (logits, posteriors, column names) and
learn_genotypes (learnt betas, last posteriors, per-iteration logits / additions from staged_genotype_learning).
The GPU box has no /root/reference: tests there read only these files.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

from reference_loader import load_reference  # noqa: E402
from demuxalot_b200.synthetic import make_dataset  # noqa: E402

CASES = {
    # name: (generator kwargs, run kwargs)
    'g4_dp25': (dict(n_genotypes=4, n_snps=150, n_barcodes=40, rows_per_barcode=120, seed=101,
                     doublet_fraction=0.25, tiny_error_fraction=0.0005),
                dict(doublet_prior=0.25, p_genotype_clip=0.01, n_iterations=5, prior=False, default_prior=1.0)),
    'g7_dp0_prior': (dict(n_genotypes=7, n_snps=260, n_barcodes=50, rows_per_barcode=90, seed=102,
                          shuffle_variants=True, spare_capacity=23, third_allele_fraction=0.1),
                     dict(doublet_prior=0.0, p_genotype_clip=0.01, n_iterations=3, prior=True, default_prior=1.0)),
    'g12_dp35': (dict(n_genotypes=12, n_snps=500, n_barcodes=64, rows_per_barcode=200, seed=103,
                      shuffle_variants=True, unknown_genotype_fraction=0.3),
                 dict(doublet_prior=0.35, p_genotype_clip=0.02, n_iterations=4, prior=True, default_prior=0.5)),
    'g33_dp35_lowdepth': (dict(n_genotypes=33, n_snps=800, n_barcodes=48, rows_per_barcode=40, seed=104,
                               empty_barcode_fraction=0.1),
                          dict(doublet_prior=0.35, p_genotype_clip=0.01, n_iterations=3, prior=False,
                               default_prior=1.0)),
}


def flatten_inputs(ds, run) -> dict:
    out = {}
    keys = list(ds.genotypes.var2varid.items())
    out['var_chrom'] = np.array([k[0] for k, _ in keys])
    out['var_pos'] = np.array([k[1] for k, _ in keys], dtype=np.int64)
    out['var_base'] = np.array([k[2] for k, _ in keys])
    out['var_id'] = np.array([v for _, v in keys], dtype=np.int32)
    out['raw_betas'] = np.array(ds.genotypes.get_betas())
    out['genotype_names'] = np.array(ds.genotypes.genotype_names)
    out['default_prior'] = np.float64(run['default_prior'])
    out['barcodes'] = np.array(ds.barcode_handler.ordered_barcodes)
    out['chromosomes'] = np.array(list(ds.calls))
    for chrom, calls in ds.calls.items():
        out[f'mol__{chrom}'] = calls.molecules  # over-allocated arrays are stored as they are
        out[f'calls__{chrom}'] = calls.snp_calls
        out[f'n__{chrom}'] = np.array([calls.n_molecules, calls.n_snp_calls], dtype=np.int64)
    return out


def make_case(ref, name: str, ds, run: dict) -> None:
    """Runs the reference on one dataset (anything with .genotypes / .calls / .barcode_handler) and stores the fixture."""
    R = ref.Demultiplexer
    ds.genotypes.default_prior = run['default_prior']
    dp, clip, n_it = run['doublet_prior'], run['p_genotype_clip'], run['n_iterations']
    fx = flatten_inputs(ds, run)
    fx['doublet_prior'], fx['p_genotype_clip'], fx['n_iterations'] = np.float64(dp), np.float64(clip), np.int64(n_it)

    v2s, betas_learn, mol, rows = R.pack_calls(ds.calls, ds.genotypes, add_data_prior=True)
    _, betas_predict, _, _ = R.pack_calls(ds.calls, ds.genotypes, add_data_prior=False)
    fx['variant2snp'] = v2s
    fx['betas_reg_learn'], fx['betas_reg_predict'] = betas_learn, betas_predict
    fx['mol_variant_id'] = mol['variant_id']
    for field in ('variant_id', 'snp_id', 'compressed_cb', 'p_base_wrong', 'barcode_variant_count'):
        fx[f'rows_{field}'] = np.asarray(rows[field])
    fx['table_predict'] = R._compute_probs_from_betas(v2s, betas_predict, p_genotype_clip=clip)

    logits_df, probs_df = R.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler,
                                               p_genotype_clip=clip, doublet_prior=dp)
    fx['predict_logits'], fx['predict_post'] = logits_df.values, probs_df.values
    fx['columns'] = np.array(list(logits_df.columns))

    prior = None
    if run['prior']:
        rng = np.random.default_rng(7)
        prior = (rng.normal(size=logits_df.shape) * 2).astype(np.float32).astype(np.float64)
        prior[rng.random(len(prior)) < 0.2, 0] += 100.  # "labelled" barcodes as in tests/test_synthetic.py:225
        fx['prior_logits'] = prior
    stages = list(R.staged_genotype_learning(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=n_it,
                                             p_genotype_clip=clip, doublet_prior=dp, barcode_prior_logits=prior))
    fx['stage_logits'] = np.stack([dbg['barcode_logits'] for _, dbg in stages])
    fx['stage_post'] = np.stack([df.values for df, _ in stages])
    fx['stage_addition'] = np.stack([dbg['genotype_addition'] for _, dbg in stages])
    learnt, post_df = R.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=n_it,
                                        p_genotype_clip=clip, doublet_prior=dp, barcode_prior_logits=prior)
    fx['learnt_betas'], fx['learn_post'] = np.array(learnt.get_betas()), post_df.values
    path = HERE / f'{name}.npz'
    np.savez_compressed(path, **fx)
    print(f'{name}: V={len(v2s)} rows={len(rows)} matched={len(mol)} C={logits_df.shape[1]} '
          f'-> {path.name} ({path.stat().st_size / 1024:.0f} KiB)')


def main() -> None:
    ref = load_reference()
    assert ref is not None, '/root/reference is required to generate golden vectors'
    R = ref.Demultiplexer
    for name, (gen, run) in CASES.items():
        make_case(ref, name, make_dataset(**gen), run)

    # known-answer table for the doublet prior (reference tests/test_utils.py:34-40)
    kat = {}
    for g in (2, 3, 10, 32):
        for dp in (0., 0.25, 0.35, 0.5):
            kat[f'pen_{g}_{dp}'] = R._doublet_penalties(g, dp)
    np.savez_compressed(HERE / 'doublet_penalties.npz', **kat)

    # a betas parquet written by the reference (layout to keep: genotypes.py:336-358)
    ds = make_dataset(n_genotypes=3, n_snps=20, n_barcodes=8, rows_per_barcode=5, seed=105)
    ref_geno = ref.ProbabilisticGenotypes(ds.genotypes.genotype_names)
    ref_geno.var2varid = dict(ds.genotypes.var2varid)
    ref_geno.variant_betas = ds.genotypes.variant_betas.copy()
    ref_geno.save_betas(HERE / 'reference_betas.parquet')
    print('wrote doublet_penalties.npz, reference_betas.parquet')


if __name__ == '__main__':
    main()
