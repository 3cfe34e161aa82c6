"""
Device-side synthetic generator (bench / test tooling, demuxalot_b200/synthetic_device.py + csrc/synth.cu): its numpy
mirror is self-consistent on CPU; on the GPU the CUDA generator reproduces the mirror byte for byte and the row builder
fed from device-resident records gives the rows the oracle builds from the mirrored host records.
"""
import numpy as np
import pytest

import oracle
from golden_io import bits


def _small():
    from demuxalot_b200.synthetic_device import make_device_dataset
    return make_device_dataset(n_genotypes=12, n_snps=4000, n_barcodes=60, rows_per_barcode=200, seed=5)


def test_host_mirror_is_keyed_by_barcode():
    """A barcode's calls do not depend on which other barcodes are generated with it (any GPU count sees one data
    set), and their relative order is fixed."""
    ds = _small()
    everything = ds.host_calls(np.arange(60), relabel=False)['chr1']
    ids = np.array([3, 7, 11, 40, 59])
    subset = ds.host_calls(ids, relabel=False)['chr1']
    all_calls = everything.snp_calls[:everything.n_snp_calls]
    all_cb = everything.molecules['compressed_cb'][all_calls['molecule_index']]
    keep = np.isin(all_cb, ids)
    sub_calls = subset.snp_calls[:subset.n_snp_calls]
    for field in ('snp_position', 'base_index', 'p_base_wrong'):
        assert np.array_equal(all_calls[field][keep], sub_calls[field]), field
    assert np.array_equal(all_cb[keep], subset.molecules['compressed_cb'][sub_calls['molecule_index']])
    # a sane data set: most calls match a variant, groups hold several molecules, donors are recoverable
    calls = ds.host_calls(np.arange(60))
    _, _, mol, rows = oracle.OracleDemultiplexer.pack_calls(calls, ds.genotypes, True)
    assert 0.9 < len(mol['variant_id']) / everything.n_snp_calls < 1.0
    assert 1.3 < len(mol['variant_id']) / len(rows['variant_id']) < 2.2
    _, post = oracle.OracleDemultiplexer.predict_posteriors(calls, ds.genotypes, ds.barcode_handler, doublet_prior=0.)
    singlet = (ds.barcode_donors[:, 1] < 0) & (ds.groups_per_barcode > 50)
    assert (post.values.argmax(axis=1)[singlet] == ds.barcode_donors[singlet, 0]).all()


def test_lanes_share_donors():
    from demuxalot_b200.synthetic_device import make_device_dataset
    a = make_device_dataset(n_genotypes=6, n_snps=500, n_barcodes=20, rows_per_barcode=50, seed=9, calls_seed=0)
    b = make_device_dataset(n_genotypes=6, n_snps=500, n_barcodes=20, rows_per_barcode=50, seed=9, calls_seed=1)
    assert a.genotypes.var2varid == b.genotypes.var2varid
    assert np.array_equal(a.genotypes.get_betas(), b.genotypes.get_betas())
    ca, cb = a.host_calls(np.arange(20))['chr1'], b.host_calls(np.arange(20))['chr1']
    assert ca.n_snp_calls != cb.n_snp_calls or not np.array_equal(ca.snp_calls, cb.snp_calls)


@pytest.mark.gpu
def test_device_generator_equals_host_mirror_and_feeds_the_row_builder(native_lib):
    import torch
    from demuxalot_b200 import Demultiplexer as D
    ds = _small()
    ids = np.array([3, 7, 11, 40, 59])
    part = ds.device_calls(ids, 'cuda')
    host = ds.host_calls(ids, relabel=False)['chr1']
    assert part['n_calls'] == host.n_snp_calls
    assert np.array_equal(part['records'].cpu().numpy().reshape(-1), host.snp_calls[:host.n_snp_calls].view(np.uint8))
    assert np.array_equal(part['molecule_cb'].cpu().numpy(), host.molecules['compressed_cb'][:host.n_molecules])
    # device-resident records -> rows: identical to the rows built from the mirrored host records and to the oracle's
    everything = ds.device_calls(np.arange(60), 'cuda')
    calls = ds.host_calls(np.arange(60))
    from_device = D._pack_device(None, ds.genotypes, 60, add_data_prior=True, device_parts=[everything])
    from_host = D._pack_device(calls, ds.genotypes, 60, add_data_prior=True)
    _, obetas, _, orows = oracle.OracleDemultiplexer.pack_calls(calls, ds.genotypes, True)
    for pack in (from_device, from_host):
        assert np.array_equal(pack.csc_variant.cpu().numpy(), orows['variant_id'])
        assert np.array_equal(pack.csc_cb.cpu().numpy(), orows['compressed_cb'])
        assert np.array_equal(bits(pack.csc_e.cpu().numpy()), bits(orows['p_base_wrong']))
        assert np.array_equal(bits(pack.betas.cpu().numpy()), bits(obetas))
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_em_with_detected_snvs_shape_vs_oracle(native_lib):
    """BASELINE config #3 in miniature: most variants are "detected SNVs" (position known, genotype unknown: zero betas,
    snp_detection.py:230-242), 10 EM iterations with doublet columns -- learnt betas within 1e-5 relative of the oracle,
    posteriors within the logit-implied bound, on device-generated calls fed through the public API."""
    from demuxalot_b200 import Demultiplexer as D
    from demuxalot_b200.synthetic_device import make_device_dataset
    ds = make_device_dataset(n_genotypes=32, n_snps=6000, n_barcodes=96, rows_per_barcode=400, seed=20260002,
                             unknown_genotype_fraction=0.78)
    calls = ds.host_calls(np.arange(96))
    assert (np.array(ds.genotypes.get_betas()).sum(axis=1) == 0).mean() > 0.7
    learnt, post = D.learn_genotypes(calls, ds.genotypes, ds.barcode_handler, n_iterations=10, doublet_prior=0.35)
    want, want_post = oracle.OracleDemultiplexer.learn_genotypes(calls, ds.genotypes, ds.barcode_handler,
                                                                 n_iterations=10, doublet_prior=0.35)
    got, ref = np.array(learnt.get_betas(), np.float64), np.array(want.get_betas(), np.float64)
    assert (np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)).max() <= 1e-5
    assert np.abs(post.values - want_post.values).max() <= 1e-5
    assert (post.values.argmax(axis=1) == want_post.values.argmax(axis=1)).all()
