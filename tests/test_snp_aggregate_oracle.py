"""
Pins the oracle's restatement of the `aggregate_on_snps = True` likelihood (demux.py:204-244,
oracle.snp_aggregated_logits) against tests/golden/aggregate_on_snps.npz, written by the unmodified reference
(tests/golden/make_golden_aggregate.py), and -- where /root/reference is mounted -- against the live reference,
bit for bit.  The group structure is integer work and must be exact everywhere; logits / posteriors go through
numpy's float32 log / exp and float64 exp / log1p, bit-exact on the machine that wrote the fixture.
"""
import numpy as np
import pytest

import oracle
from golden_io import CASES, GOLDEN_DIR, bits, load_case
from reference_loader import load_reference, reference_available


@pytest.fixture(scope='module')
def golden():
    return dict(np.load(GOLDEN_DIR / 'aggregate_on_snps.npz', allow_pickle=False))


@pytest.fixture()
def aggregated_oracle():
    oracle.OracleDemultiplexer.aggregate_on_snps = True
    try:
        yield oracle.OracleDemultiplexer
    finally:
        oracle.OracleDemultiplexer.aggregate_on_snps = False


@pytest.mark.parametrize('name', CASES)
def test_group_structure_exact(name, golden):
    case = load_case(name)
    _, _, mol, _ = oracle.OracleDemultiplexer.pack_calls(case.calls, case.genotypes, False)
    group_of_call, counts, group_barcode = oracle.snp_groups(mol['compressed_cb'], mol['snp_id'])
    assert np.array_equal(counts, golden[f'{name}__group_counts'])
    assert np.array_equal(group_barcode, golden[f'{name}__group_barcode'])
    first_call = np.unique(group_of_call, return_index=True)[1]
    assert np.array_equal(mol['snp_id'][first_call], golden[f'{name}__group_snp'])


@pytest.mark.parametrize('name', CASES)
def test_oracle_against_golden(name, golden, aggregated_oracle):
    case = load_case(name)
    O = aggregated_oracle
    kwargs = dict(p_genotype_clip=case.p_genotype_clip, doublet_prior=case.doublet_prior)
    logits_df, probs_df = O.predict_posteriors(case.calls, case.genotypes, case.barcode_handler, **kwargs)
    assert logits_df.values.dtype == np.float64 and probs_df.values.dtype == np.float64
    assert list(logits_df.columns) == [str(c) for c in golden[f'{name}__columns']]
    np.testing.assert_allclose(logits_df.values, golden[f'{name}__predict_logits'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(probs_df.values, golden[f'{name}__predict_post'], rtol=0, atol=5e-6)
    learn = dict(kwargs, n_iterations=case.n_iterations, barcode_prior_logits=case.prior_logits)
    stages = list(O.staged_genotype_learning(case.calls, case.genotypes, case.barcode_handler, **learn))
    for it, (post_df, dbg) in enumerate(stages):
        np.testing.assert_allclose(dbg['barcode_logits'], golden[f'{name}__stage_logits'][it], rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(post_df.values, golden[f'{name}__stage_post'][it], rtol=0, atol=5e-5)
        np.testing.assert_allclose(dbg['genotype_addition'], golden[f'{name}__stage_addition'][it], rtol=1e-4, atol=1e-6)
    learnt, post_df = O.learn_genotypes(case.calls, case.genotypes, case.barcode_handler, **learn)
    np.testing.assert_allclose(learnt.get_betas(), golden[f'{name}__learnt_betas'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(post_df.values, golden[f'{name}__learn_post'], rtol=0, atol=5e-5)


@pytest.mark.skipif(not reference_available(), reason='/root/reference not mounted')
@pytest.mark.parametrize('shape', [
    dict(n_genotypes=5, n_snps=300, n_barcodes=30, rows_per_barcode=100, seed=21, shuffle_variants=True),
    dict(n_genotypes=16, n_snps=2000, n_barcodes=60, rows_per_barcode=250, seed=22, third_allele_fraction=0.2),
])
def test_oracle_equals_live_reference(shape, aggregated_oracle):
    from demuxalot_b200.synthetic import make_dataset
    ref = load_reference()
    ds = make_dataset(**shape)
    R, O = ref.Demultiplexer, aggregated_oracle
    R.aggregate_on_snps = True
    try:
        for dp in (0., 0.35):
            rl, rp = R.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=dp)
            ol, op = O.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=dp)
            assert list(rl.columns) == list(ol.columns) and rl.index.equals(ol.index)
            assert np.array_equal(bits(rl.values), bits(ol.values))
            assert np.array_equal(bits(rp.values), bits(op.values))
            prior = np.random.default_rng(0).normal(size=rl.shape) * 3
            rg, rpost = R.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=dp,
                                          n_iterations=3, barcode_prior_logits=prior)
            og, opost = O.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=dp,
                                          n_iterations=3, barcode_prior_logits=prior)
            assert np.array_equal(bits(np.array(rg.get_betas())), bits(np.array(og.get_betas())))
            assert np.array_equal(bits(rpost.values), bits(opost.values))
    finally:
        R.aggregate_on_snps = False
