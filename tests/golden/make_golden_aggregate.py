"""
Generates tests/golden/aggregate_on_snps.npz: outputs of the UNMODIFIED reference with the class flag
`Demultiplexer.aggregate_on_snps = True` (demux.py:31,204-244) on the inputs of the existing fixtures
(tests/golden/<case>.npz, rebuilt through golden_io.load_case).  Run in the build container:

    python tests/golden/make_golden_aggregate.py

Per case: predict_posteriors logits / posteriors (float64 in this branch), per-iteration logits, posteriors and
additions of staged_genotype_learning, the learnt betas and last posteriors of learn_genotypes, and the
(barcode, SNP) group structure (number of groups, molecules per group, barcode of each group).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

from golden_io import CASES, load_case  # noqa: E402
from reference_loader import load_reference  # noqa: E402


def main() -> None:
    ref = load_reference()
    assert ref is not None, '/root/reference is required to generate golden vectors'
    R = ref.Demultiplexer
    out = {}
    R.aggregate_on_snps = True
    try:
        for name in CASES:
            case = load_case(name)
            kwargs = dict(p_genotype_clip=case.p_genotype_clip, doublet_prior=case.doublet_prior)
            logits_df, probs_df = R.predict_posteriors(case.calls, case.genotypes, case.barcode_handler, **kwargs)
            assert logits_df.values.dtype == np.float64 and probs_df.values.dtype == np.float64
            out[f'{name}__predict_logits'], out[f'{name}__predict_post'] = logits_df.values, probs_df.values
            out[f'{name}__columns'] = np.array(list(logits_df.columns))
            learn = dict(kwargs, n_iterations=case.n_iterations, barcode_prior_logits=case.prior_logits)
            stages = list(R.staged_genotype_learning(case.calls, case.genotypes, case.barcode_handler, **learn))
            out[f'{name}__stage_logits'] = np.stack([dbg['barcode_logits'] for _, dbg in stages])
            out[f'{name}__stage_post'] = np.stack([df.values for df, _ in stages])
            out[f'{name}__stage_addition'] = np.stack([dbg['genotype_addition'] for _, dbg in stages])
            learnt, post_df = R.learn_genotypes(case.calls, case.genotypes, case.barcode_handler, **learn)
            out[f'{name}__learnt_betas'], out[f'{name}__learn_post'] = np.array(learnt.get_betas()), post_df.values
            # group structure as the reference's FeatureLookup sees it (utils.py:207-265)
            _, _, mol, _ = R.pack_calls(case.calls, case.genotypes, add_data_prior=False)
            lookup = ref.utils.FeatureLookup(mol['compressed_cb'], mol['snp_id'])
            _ids, counts = lookup.compress(mol['compressed_cb'], mol['snp_id'])
            group_barcode, group_snp = lookup.lookup_for_individual_features()
            out[f'{name}__group_counts'] = counts.astype(np.int64)
            out[f'{name}__group_barcode'] = np.asarray(group_barcode).astype(np.int64)
            out[f'{name}__group_snp'] = np.asarray(group_snp).astype(np.int64)
            print(f'{name}: groups={len(counts)} C={logits_df.shape[1]} iterations={len(stages)}')
    finally:
        R.aggregate_on_snps = False
    path = HERE / 'aggregate_on_snps.npz'
    np.savez_compressed(path, **out)
    print(f'wrote {path.name} ({path.stat().st_size / 1024:.0f} KiB)')


if __name__ == '__main__':
    main()
