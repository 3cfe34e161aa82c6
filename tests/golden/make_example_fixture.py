"""
Config #1 fixture: the reference's bundled example (examples/example_data: BAM + VCF + barcodes, 4 donors) run
through the UNMODIFIED reference -- `add_vcf`, `count_snps`, `learn_genotypes(doublet_prior=0.25)` exactly as
examples/1-plain_demultiplexing.py does -- with tests/pysam_shim.py standing in for pysam.

    python tests/golden/make_example_fixture.py

The full count_snps output is 4.0 M calls (52 MB), too large to commit, so the stored fixture holds the calls of the
first 48 barcodes (`example_data_48bc.npz`, same layout as the synthetic fixtures: inputs + reference outputs for
pack_calls / predict_posteriors / learn_genotypes on that subset), plus `example_data_summary.json` with the
sizes and per-donor call counts of the full run, and the genotypes our own `add_vcf` must reproduce bit for bit.
"""
import json
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

import pysam_shim  # noqa: E402

sys.modules['pysam'] = pysam_shim
sys.path.insert(0, '/root/reference')
import demuxalot as ref  # noqa: E402  (the reference)

from make_golden import make_case  # noqa: E402
from bench import slice_barcodes  # noqa: E402

DATA = Path('/root/reference/examples/example_data')
N_BARCODES = 48


def main():
    donors = ['Donor01', 'Donor02', 'Donor03', 'Donor04']
    genotypes = ref.ProbabilisticGenotypes(genotype_names=donors)
    genotypes.add_vcf(str(DATA / 'test_genotypes.vcf'))
    barcode_handler = ref.BarcodeHandler.from_file(str(DATA / 'test_barcodes.csv'))
    calls = ref.count_snps(bamfile_location=str(DATA / 'test_bamfile.bam'),
                           chromosome2positions=genotypes.get_chromosome2positions(),
                           barcode_handler=barcode_handler, joblib_n_jobs=1, joblib_verbosity=0)
    learnt, post = ref.Demultiplexer.learn_genotypes(calls, genotypes=genotypes, barcode_handler=barcode_handler,
                                                     doublet_prior=0.25)
    _, _, mol, rows = ref.Demultiplexer.pack_calls(calls, genotypes, add_data_prior=True)
    best = post.values.argmax(axis=1)
    summary = dict(
        n_variants=genotypes.n_variants, n_barcodes=barcode_handler.n_barcodes,
        chromosomes={c: dict(n_molecules=int(v.n_molecules), n_snp_calls=int(v.n_snp_calls)) for c, v in calls.items()},
        n_matched_calls=int(len(mol)), n_rows=int(len(rows)),
        argmax_counts={str(post.columns[k]): int((best == k).sum()) for k in range(post.shape[1])},
        min_max_posterior=float(post.values.max(axis=1).min()),
    )
    (HERE / 'example_data_summary.json').write_text(json.dumps(summary, indent=1))
    print(json.dumps(summary))

    # genotypes as the reference's add_vcf builds them (pins our plain-text add_vcf)
    keys = list(genotypes.var2varid.items())
    np.savez_compressed(HERE / 'example_genotypes.npz',
                        var_chrom=np.array([k[0] for k, _ in keys]), var_pos=np.array([k[1] for k, _ in keys]),
                        var_base=np.array([k[2] for k, _ in keys]), var_id=np.array([v for _, v in keys]),
                        betas=np.array(genotypes.get_betas()))

    ds = SimpleNamespace(genotypes=genotypes, calls=calls, barcode_handler=barcode_handler)
    sub_calls, sub_handler = slice_barcodes(ds, N_BARCODES)
    sub = SimpleNamespace(genotypes=genotypes, calls=sub_calls, barcode_handler=sub_handler)
    make_case(ref, f'example_data_{N_BARCODES}bc', sub,
              dict(doublet_prior=0.25, p_genotype_clip=0.01, n_iterations=5, prior=False, default_prior=1.0))


if __name__ == '__main__':
    main()
