"""
Parity of the CUDA path with `Demultiplexer.aggregate_on_snps = True` (demux.py:204-244; dmx_build_snp_groups,
dmx_snp_logits, dmx_softmax_rows_f64 through the public API) against
  (1) tests/golden/aggregate_on_snps.npz, written by the unmodified reference with the flag flipped, and
  (2) the oracle (oracle.snp_aggregated_logits) on fresh seeded inputs, including a 528-column case.

Bars: the (barcode, SNP) group structure is integer work and must be exact; logits (float64 in this branch)
within 1e-5 relative (+1e-4 absolute: a barcode's logit is a sum over its groups of terms that each carry a few
float32 ulp of the host's / device's logf); posteriors within 1e-6 plus the bound implied by the logit difference,
as in test_gpu_parity.py; learnt betas within 1e-5 relative (+1e-5 absolute on the additions).
Measured differences go to gpurun_out/parity_report_snp_aggregate.json.
"""
import json
import os
from pathlib import Path

import numpy as np
import pytest

import oracle
from golden_io import CASES, GOLDEN_DIR, bits, load_case

pytestmark = pytest.mark.gpu

REPORT = {}


@pytest.fixture(scope='module')
def golden():
    return dict(np.load(GOLDEN_DIR / 'aggregate_on_snps.npz', allow_pickle=False))


@pytest.fixture(scope='module')
def D(native_lib):
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from demuxalot_b200 import Demultiplexer
    Demultiplexer.aggregate_on_snps = True
    oracle.OracleDemultiplexer.aggregate_on_snps = True
    try:
        yield Demultiplexer
    finally:
        Demultiplexer.aggregate_on_snps = False
        oracle.OracleDemultiplexer.aggregate_on_snps = False
        out = Path(os.environ.get('GRAFT_REPO_ROOT', Path(__file__).resolve().parent.parent)) / 'gpurun_out'
        out.mkdir(exist_ok=True)
        (out / 'parity_report_snp_aggregate.json').write_text(json.dumps(REPORT, indent=1, sort_keys=True))


def check(tag, got_logits, want_logits, got_post, want_post):
    got_logits, want_logits = np.asarray(got_logits), np.asarray(want_logits)
    assert got_logits.dtype == np.float64 and np.asarray(got_post).dtype == np.float64, 'this branch is float64'
    assert got_logits.shape == want_logits.shape
    dlogit = np.abs(got_logits - want_logits)
    dpost = np.abs(np.asarray(got_post) - np.asarray(want_post))
    REPORT[tag] = dict(logit_abs_max=float(dlogit.max(initial=0)),
                       logit_rel_max=float((dlogit / np.maximum(np.abs(want_logits), 1e-30)).max(initial=0)),
                       post_abs_max=float(dpost.max(initial=0)),
                       post_le_1e6_frac=float((dpost <= 1e-6).mean()) if dpost.size else 1.)
    assert (dlogit <= 1e-5 * np.abs(want_logits) + 1e-4).all(), f'{tag}: logits differ by {dlogit.max()}'
    # a softmax entry moves by at most half the spread of the logit differences within its row (<= max |d logit|)
    bound = 1e-6 + dlogit.max(axis=1)[:, None]
    assert (dpost <= bound).all(), f'{tag}: posteriors differ by {dpost.max()} beyond the logit-implied bound'
    assert np.allclose(np.asarray(got_post).sum(axis=1), 1, atol=1e-9)


@pytest.mark.parametrize('name', CASES)
def test_groups_exact_vs_reference_fixture(D, golden, name):
    case = load_case(name)
    pack = D._pack_device(case.calls, case.genotypes, case.barcode_handler.n_barcodes, add_data_prior=False)
    grouped_variant, grouped_e, group_offsets, barcode_group_offsets, n_matched, n_groups = D._snp_groups(pack)
    counts, group_barcode = golden[f'{name}__group_counts'], golden[f'{name}__group_barcode']
    assert n_groups == len(counts) and n_matched == int(counts.sum()) == len(case.fx['mol_variant_id'])
    offsets = group_offsets.cpu().numpy()[:n_groups + 1]
    assert np.array_equal(np.diff(offsets), counts) and offsets[0] == 0
    B = case.barcode_handler.n_barcodes
    assert np.array_equal(barcode_group_offsets.cpu().numpy(), np.searchsorted(group_barcode, np.arange(B + 1)))
    # calls in group order, original order inside a group: a stable sort of the matched calls by (barcode, SNP)
    _, _, mol, _ = oracle.OracleDemultiplexer.pack_calls(case.calls, case.genotypes, False)
    keys = mol['compressed_cb'].astype(np.int64) * (int(mol['snp_id'].max()) + 1) + mol['snp_id']
    order = np.argsort(keys, kind='stable')
    assert np.array_equal(grouped_variant.cpu().numpy()[:n_matched], mol['variant_id'][order])
    assert np.array_equal(bits(grouped_e.cpu().numpy()[:n_matched]), bits(mol['p_base_wrong'][order]))
    assert np.array_equal(pack.variant2snp[mol['variant_id'][order]][offsets[:-1]], golden[f'{name}__group_snp'])


@pytest.mark.parametrize('name', CASES)
def test_predict_posteriors_vs_reference_fixture(D, golden, name):
    case = load_case(name)
    logits_df, probs_df = D.predict_posteriors(case.calls, case.genotypes, case.barcode_handler,
                                               p_genotype_clip=case.p_genotype_clip, doublet_prior=case.doublet_prior)
    assert list(logits_df.columns) == [str(c) for c in golden[f'{name}__columns']]
    assert logits_df.index.name == 'BARCODE' and list(logits_df.index) == case.barcode_handler.ordered_barcodes
    check(f'predict/{name}', logits_df.values, golden[f'{name}__predict_logits'], probs_df.values,
          golden[f'{name}__predict_post'])


@pytest.mark.parametrize('name', CASES)
def test_learn_genotypes_vs_reference_fixture(D, golden, name):
    case = load_case(name)
    kwargs = dict(n_iterations=case.n_iterations, p_genotype_clip=case.p_genotype_clip,
                  doublet_prior=case.doublet_prior, barcode_prior_logits=case.prior_logits)
    stages = list(D.staged_genotype_learning(case.calls, case.genotypes, case.barcode_handler, **kwargs))
    assert len(stages) == case.n_iterations
    for it, (post_df, dbg) in enumerate(stages):
        check(f'stage{it}/{name}', dbg['barcode_logits'], golden[f'{name}__stage_logits'][it], post_df.values,
              golden[f'{name}__stage_post'][it])
        np.testing.assert_allclose(dbg['genotype_addition'], golden[f'{name}__stage_addition'][it], rtol=1e-5, atol=1e-5)
    learnt, post_df = D.learn_genotypes(case.calls, case.genotypes, case.barcode_handler, **kwargs)
    want = golden[f'{name}__learnt_betas']
    got = np.asarray(learnt.get_betas())
    REPORT[f'learn/{name}'] = dict(betas_rel_max=float((np.abs(got - want) / np.maximum(np.abs(want), 1e-3)).max()))
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
    dpost = np.abs(post_df.values - golden[f'{name}__learn_post'])
    assert dpost.max() <= 1e-4, f'learn/{name}: last posteriors differ by {dpost.max()}'


@pytest.mark.parametrize('shape,dp', [
    (dict(n_genotypes=32, n_snps=3000, n_barcodes=96, rows_per_barcode=400, seed=31), 0.35),  # 528 columns
    (dict(n_genotypes=3, n_snps=50, n_barcodes=300, rows_per_barcode=30, seed=32, empty_barcode_fraction=0.2), 0.35),
    (dict(n_genotypes=40, n_snps=1500, n_barcodes=64, rows_per_barcode=2500, seed=33), 0.),  # deep groups
    (dict(n_genotypes=9, n_snps=400, n_barcodes=50, rows_per_barcode=150, seed=34, third_allele_fraction=0.3,
          shuffle_variants=True), 0.2),
])
def test_against_oracle_on_fresh_inputs(D, shape, dp):
    from demuxalot_b200.synthetic import make_dataset
    ds = make_dataset(**shape)
    O = oracle.OracleDemultiplexer
    want_logits, want_post = O.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=dp)
    got_logits, got_post = D.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=dp)
    assert list(got_logits.columns) == list(want_logits.columns)
    check(f'oracle/G{shape["n_genotypes"]}_dp{dp}', got_logits.values, want_logits.values, got_post.values,
          want_post.values)
    prior = np.random.default_rng(5).normal(size=want_logits.shape) * 2
    want_geno, want_last = O.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=dp,
                                             n_iterations=3, barcode_prior_logits=prior)
    got_geno, got_last = D.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=dp,
                                           n_iterations=3, barcode_prior_logits=prior)
    np.testing.assert_allclose(np.asarray(got_geno.get_betas()), np.asarray(want_geno.get_betas()), rtol=1e-5, atol=1e-5)
    assert np.abs(got_last.values - want_last.values).max() <= 1e-4


def test_runs_are_deterministic(D):
    case = load_case('g12_dp35')
    a, _ = D.predict_posteriors(case.calls, case.genotypes, case.barcode_handler, doublet_prior=case.doublet_prior)
    b, _ = D.predict_posteriors(case.calls, case.genotypes, case.barcode_handler, doublet_prior=case.doublet_prior)
    assert np.array_equal(bits(a.values), bits(b.values))


def test_flag_off_selects_the_float32_path(D):
    D.aggregate_on_snps = False  # the module fixture holds it at True
    try:
        case = load_case('g4_dp25')
        logits_df, _ = D.predict_posteriors(case.calls, case.genotypes, case.barcode_handler,
                                            doublet_prior=case.doublet_prior)
        assert logits_df.values.dtype == np.float32
        np.testing.assert_allclose(logits_df.values, case.fx['predict_logits'], rtol=1e-5)
    finally:
        D.aggregate_on_snps = True
