"""
Pickles the reference's OWN objects -- demuxalot.ProbabilisticGenotypes, demuxalot.BarcodeHandler and
demuxalot.snp_counter.CompressedSNPCalls instances built with the unmodified reference (/root/reference, pysam stubbed) --
together with what the reference computes from them, into tests/golden/ref_objects.pkl:

    python tests/golden/make_ref_objects.py

The GPU box has no reference package: tests/test_gpu_reference_objects.py registers attribute-compatible stand-in
classes under the reference's module paths before unpickling, so that the restored objects carry exactly the
reference's state (`var2varid`, `variant_betas`, `genotype_names`, `default_prior`; `barcode2index`,
`ordered_barcodes`; `molecules`, `snp_calls`, `n_molecules`, `n_snp_calls`) and are fed to the CUDA path as they are.
"""
from __future__ import annotations

import pickle
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

from reference_loader import load_reference  # noqa: E402
from demuxalot_b200.synthetic import make_dataset  # noqa: E402


def main():
    ref = load_reference()
    assert ref is not None, 'needs /root/reference'
    from demuxalot.snp_counter import CompressedSNPCalls as RefCalls
    ds = make_dataset(n_genotypes=9, n_snps=300, n_barcodes=48, rows_per_barcode=110, seed=909, shuffle_variants=True,
                      third_allele_fraction=0.05)
    genotypes = ref.ProbabilisticGenotypes(list(ds.genotypes.genotype_names), default_prior=0.7)
    # direct attribute injection, as the reference's own tests build their genotypes (tests/test_synthetic.py:79-103)
    genotypes.var2varid = dict(ds.genotypes.var2varid)
    genotypes.variant_betas = np.array(ds.genotypes.variant_betas)
    handler = ref.BarcodeHandler(list(ds.barcode_handler.ordered_barcodes))
    calls = {}
    for chrom, c in ds.calls.items():
        rc = RefCalls()
        rc.molecules, rc.n_molecules = c.molecules[:c.n_molecules].copy(), c.n_molecules
        rc.snp_calls, rc.n_snp_calls = c.snp_calls[:c.n_snp_calls].copy(), c.n_snp_calls
        calls[chrom] = rc
    logits, post = ref.Demultiplexer.predict_posteriors(calls, genotypes, handler, doublet_prior=0.35)
    learnt, learnt_post = ref.Demultiplexer.learn_genotypes(calls, genotypes, handler, n_iterations=4, doublet_prior=0.35)
    payload = dict(genotypes=genotypes, barcode_handler=handler, calls=calls,
                   expected=dict(columns=list(logits.columns), index=list(logits.index), logits=logits.values,
                                 posteriors=post.values, learnt_betas=np.array(learnt.get_betas()),
                                 learnt_posteriors=learnt_post.values))
    with open(HERE / 'ref_objects.pkl', 'wb') as f:
        pickle.dump(payload, f, protocol=4)
    print('wrote', HERE / 'ref_objects.pkl', (HERE / 'ref_objects.pkl').stat().st_size, 'bytes')


if __name__ == '__main__':
    main()
