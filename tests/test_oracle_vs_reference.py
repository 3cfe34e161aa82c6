"""
Live pin of the oracle: runs the unmodified reference from /root/reference next to the oracle on fresh seeded
inputs and requires BIT-EXACT agreement (same numpy on the same CPU).  Skipped where /root/reference is absent
(the GPU box); tests/test_oracle_golden.py covers that case with the committed fixtures.
"""
import numpy as np
import pytest

from golden_io import bits
from reference_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason='/root/reference not mounted')


@pytest.mark.parametrize('shape', [
    dict(n_genotypes=5, n_snps=300, n_barcodes=30, rows_per_barcode=100, seed=11, shuffle_variants=True),
    dict(n_genotypes=2, n_snps=40, n_barcodes=12, rows_per_barcode=300, seed=12, spare_capacity=9),
    dict(n_genotypes=16, n_snps=2000, n_barcodes=100, rows_per_barcode=250, seed=13, shuffle_variants=True,
         third_allele_fraction=0.2),
])
def test_oracle_equals_reference(shape):
    import oracle
    from demuxalot_b200.synthetic import make_dataset
    ref = load_reference()
    ds = make_dataset(**shape)
    R, O = ref.Demultiplexer, oracle.OracleDemultiplexer
    for add_prior in (True, False):
        rv2s, rbetas, rmol, rrows = R.pack_calls(ds.calls, ds.genotypes, add_data_prior=add_prior)
        ov2s, obetas, omol, orows = O.pack_calls(ds.calls, ds.genotypes, add_prior)
        assert np.array_equal(rv2s, ov2s)
        assert np.array_equal(bits(rbetas), bits(obetas))
        for f in ('variant_id', 'snp_id', 'compressed_cb', 'molecule_id', 'p_base_wrong', 'p_molecule_aligned_wrong'):
            assert np.array_equal(bits(rmol[f]), bits(omol[f])), f
        for f in ('variant_id', 'snp_id', 'compressed_cb', 'p_base_wrong', 'barcode_variant_count', 'barcode_snp_count'):
            assert np.array_equal(bits(rrows[f]), bits(orows[f])), f
    for dp in (0., 0.35):
        rl, rp = R.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=dp)
        ol, op = O.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=dp)
        assert list(rl.columns) == list(ol.columns) and rl.index.equals(ol.index)
        assert np.array_equal(bits(rl.values), bits(ol.values))
        assert np.array_equal(bits(rp.values), bits(op.values))
        prior = np.random.default_rng(0).normal(size=rl.shape) * 3
        rg, rpost = R.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=dp, n_iterations=4,
                                      barcode_prior_logits=prior)
        og, opost = O.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=dp, n_iterations=4,
                                      barcode_prior_logits=prior)
        assert np.array_equal(bits(np.array(rg.get_betas())), bits(np.array(og.get_betas())))
        assert np.array_equal(bits(rpost.values), bits(opost.values))


def test_column_sharded_oracle_is_identical():
    import oracle
    from demuxalot_b200.synthetic import make_dataset
    ds = make_dataset(n_genotypes=9, n_snps=300, n_barcodes=40, rows_per_barcode=80, seed=14)
    O = oracle.OracleDemultiplexer
    l1, _ = O.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler)
    try:
        O.n_jobs = 3
        l3, _ = O.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler)
    finally:
        O.n_jobs = 1
    assert np.array_equal(bits(l1.values), bits(l3.values))


def test_our_host_types_feed_the_reference_and_parquet_roundtrip(tmp_path):
    """Our ProbabilisticGenotypes writes a parquet the reference reads back identically, and vice versa."""
    from demuxalot_b200 import ProbabilisticGenotypes
    from demuxalot_b200.synthetic import make_dataset
    ref = load_reference()
    ds = make_dataset(n_genotypes=4, n_snps=60, n_barcodes=8, rows_per_barcode=10, seed=15, shuffle_variants=True)
    ours = ds.genotypes
    ours.save_betas(tmp_path / 'ours.parquet')
    theirs = ref.ProbabilisticGenotypes(ours.genotype_names)
    theirs.add_prior_betas(tmp_path / 'ours.parquet')
    assert set(theirs.var2varid) == set(ours.var2varid)
    for key, vid in ours.var2varid.items():
        assert np.array_equal(ours.variant_betas[vid], theirs.variant_betas[theirs.var2varid[key]])
    theirs.save_betas(tmp_path / 'theirs.parquet')
    import pyarrow.parquet as pq
    a, b = pq.read_table(tmp_path / 'ours.parquet'), pq.read_table(tmp_path / 'theirs.parquet')
    assert a.schema.equals(b.schema) and a.equals(b)
    back = ProbabilisticGenotypes(ours.genotype_names)
    back.add_prior_betas(tmp_path / 'theirs.parquet')
    for key, vid in ours.var2varid.items():
        assert np.array_equal(ours.variant_betas[vid], back.variant_betas[back.var2varid[key]])


def test_plain_text_add_vcf_matches_reference_on_example_data():
    """Our pysam-free `add_vcf` builds the same genotypes as the reference's (pysam.VariantFile) on the bundled VCF;
    the expected values were written by the reference through tests/pysam_shim.py (make_example_fixture.py)."""
    from pathlib import Path
    from demuxalot_b200 import ProbabilisticGenotypes
    from golden_io import GOLDEN_DIR
    vcf = Path('/root/reference/examples/example_data/test_genotypes.vcf')
    want = np.load(GOLDEN_DIR / 'example_genotypes.npz')
    ours = ProbabilisticGenotypes(['Donor01', 'Donor02', 'Donor03', 'Donor04'])
    ours.add_vcf(vcf)
    keys = [(str(c), int(p), str(b)) for c, p, b in zip(want['var_chrom'], want['var_pos'], want['var_base'])]
    assert list(ours.var2varid.items()) == list(zip(keys, [int(v) for v in want['var_id']]))
    assert np.array_equal(bits(np.array(ours.get_betas())), bits(want['betas']))
    positions = ours.get_chromosome2positions()
    assert sorted(positions) == ['chr1', 'chr2', 'chr3'] and sum(len(p) for p in positions.values()) == 1212
