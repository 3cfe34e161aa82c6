"""
Pins the oracle (oracle/demux_oracle.py) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  Integer ids, p_base_wrong bit patterns, regularised betas and the probability
table involve no transcendental function and must be bit-exact on any CPU; logits / posteriors go through
numpy's float32 log / exp, whose last-ulp behaviour depends on the SIMD dispatch of the host CPU, so they are
bit-exact on the machine that wrote the fixtures and within a few ulp elsewhere.
"""
import numpy as np
import pytest

import oracle
from golden_io import CASES, GOLDEN_DIR, bits, load_case

LOGIT_RTOL = 2e-6  # float32 logits: allows a last-ulp difference of the host's float32 log on another CPU
POST_ATOL = 5e-5   # a 1-ulp float32 logit flip (|logit| ~ 1e2..1e3) moves an unsaturated posterior by up to this


@pytest.mark.parametrize('name', CASES)
def test_pack_calls_bit_exact(name):
    case = load_case(name)
    fx = case.fx
    v2s, betas_learn, mol, rows = oracle.OracleDemultiplexer.pack_calls(case.calls, case.genotypes, True)
    _, betas_predict, _, _ = oracle.OracleDemultiplexer.pack_calls(case.calls, case.genotypes, False)
    assert np.array_equal(v2s, fx['variant2snp'])
    assert np.array_equal(mol['variant_id'], fx['mol_variant_id'])
    for field in ('variant_id', 'snp_id', 'compressed_cb', 'barcode_variant_count'):
        assert np.array_equal(rows[field], fx[f'rows_{field}']), field
        assert rows[field].dtype == fx[f'rows_{field}'].dtype, field
    assert np.array_equal(bits(rows['p_base_wrong']), bits(fx['rows_p_base_wrong']))
    assert np.array_equal(bits(betas_learn), bits(fx['betas_reg_learn']))
    assert np.array_equal(bits(betas_predict), bits(fx['betas_reg_predict']))
    table = oracle.probs_from_betas(v2s, betas_predict, case.p_genotype_clip)
    assert np.array_equal(bits(table), bits(fx['table_predict']))


@pytest.mark.parametrize('name', CASES)
def test_predict_posteriors(name):
    case = load_case(name)
    fx = case.fx
    logits_df, probs_df = oracle.OracleDemultiplexer.predict_posteriors(
        case.calls, case.genotypes, case.barcode_handler, p_genotype_clip=case.p_genotype_clip,
        doublet_prior=case.doublet_prior)
    assert list(logits_df.columns) == [str(c) for c in fx['columns']]
    assert logits_df.index.name == 'BARCODE' and probs_df.index.name == 'BARCODE'
    assert logits_df.values.dtype == np.float32 and probs_df.values.dtype == np.float32
    np.testing.assert_allclose(logits_df.values, fx['predict_logits'], rtol=LOGIT_RTOL, atol=0)
    np.testing.assert_allclose(probs_df.values, fx['predict_post'], rtol=0, atol=POST_ATOL)


@pytest.mark.parametrize('name', CASES)
def test_learn_genotypes(name):
    case = load_case(name)
    fx = case.fx
    kwargs = dict(n_iterations=case.n_iterations, p_genotype_clip=case.p_genotype_clip,
                  doublet_prior=case.doublet_prior, barcode_prior_logits=case.prior_logits)
    stages = list(oracle.OracleDemultiplexer.staged_genotype_learning(
        case.calls, case.genotypes, case.barcode_handler, **kwargs))
    assert len(stages) == case.n_iterations
    for it, (post_df, dbg) in enumerate(stages):
        assert set(dbg) == {'barcode_logits', 'genotype_prior', 'genotype_addition'}
        np.testing.assert_allclose(dbg['barcode_logits'], fx['stage_logits'][it], rtol=LOGIT_RTOL, atol=1e-5)
        np.testing.assert_allclose(post_df.values, fx['stage_post'][it], rtol=0, atol=POST_ATOL)
        np.testing.assert_allclose(dbg['genotype_addition'], fx['stage_addition'][it], rtol=1e-4, atol=1e-6)
    learnt, post_df = oracle.OracleDemultiplexer.learn_genotypes(case.calls, case.genotypes, case.barcode_handler, **kwargs)
    np.testing.assert_allclose(learnt.get_betas(), fx['learnt_betas'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(post_df.values, fx['learn_post'], rtol=0, atol=POST_ATOL)
    assert post_df.index.name is None


def test_doublet_penalties_known_answers():
    kat = np.load(GOLDEN_DIR / 'doublet_penalties.npz')
    for key in kat.files:
        _, g, dp = key.split('_')
        got = oracle.doublet_penalties(int(g), float(dp))
        assert np.array_equal(bits(got), bits(kat[key])), key
    # the reference's own known-answer test (tests/test_utils.py:34-40): singlet prior mass is 1 - dp
    for g in (2, 3, 10):
        for dp in (0., 0.25, 0.5):
            prior = oracle.softmax_rows(oracle.doublet_penalties(g, dp)[None, :])[0]
            assert np.allclose(prior[:g].sum(), 1 - dp)


def test_softmax_matches_scipy():
    from scipy.special import softmax
    rng = np.random.default_rng(0)
    x = (rng.normal(size=(50, 78)) * 30).astype(np.float32)
    assert np.array_equal(bits(oracle.softmax_rows(x)), bits(softmax(x, axis=-1)))


def test_unknown_chromosome_raises():
    case = load_case('g4_dp25')
    calls = dict(case.calls)
    first = next(iter(calls))
    calls['chrUnknown'] = calls[first]
    with pytest.raises(AssertionError):
        oracle.OracleDemultiplexer.pack_calls(calls, case.genotypes, False)
