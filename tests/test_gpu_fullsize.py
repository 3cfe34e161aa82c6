"""
BASELINE-size checks (config #2: 32 donors, 528 columns, 657k variants, 10k barcodes, ~20M rows, 31M calls) that
do not need a full-size oracle run: structural invariants of the builder, a barcode-slice comparison against the
oracle (the E-step is barcode-local, so the first barcodes of the full run must match an oracle run on just their
calls), shard consistency, and additivity of the M-step over barcode shards.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def full(native_lib):
    import torch
    assert torch.cuda.is_available()
    from demuxalot_b200 import Demultiplexer
    from demuxalot_b200.synthetic import make_config
    ds = make_config('pbmc_32')
    B = ds.barcode_handler.n_barcodes
    pack = Demultiplexer._pack_device(ds.calls, ds.genotypes, B, add_data_prior=True)
    return ds, pack, Demultiplexer


def test_builder_invariants_at_full_size(full):
    import torch
    ds, pack, D = full
    assert pack.n_calls == ds.n_calls and 0 < pack.n_rows <= pack.n_matched <= pack.n_calls
    key = pack.csc_variant.to(torch.int64) * pack.n_barcodes + pack.csc_cb.to(torch.int64)
    assert bool((key[1:] > key[:-1]).all()), 'rows must be strictly ascending in (variant, barcode)'
    assert int(pack.csc_count.sum()) == pack.n_matched
    assert int(pack.n_mol.sum()) == pack.n_matched
    assert bool((pack.csc_count > 0).all())
    vo, bo = pack.variant_offsets, pack.barcode_offsets
    assert int(vo[0]) == 0 and int(vo[-1]) == pack.n_rows and bool((vo[1:] >= vo[:-1]).all())
    assert int(bo[0]) == 0 and int(bo[-1]) == pack.n_rows and bool((bo[1:] >= bo[:-1]).all())
    # CSR is a permutation of CSC, barcode-major and variant-ascending inside a barcode
    assert bool((torch.sort(pack.csr_row.to(torch.int64)).values == torch.arange(pack.n_rows, device=pack.device)).all())
    csr_cb = pack.csc_cb[pack.csr_row.to(torch.int64)]
    key2 = csr_cb.to(torch.int64) * (pack.n_variants + 1) + pack.csr_variant.to(torch.int64)
    assert bool((key2[1:] > key2[:-1]).all())
    # every matched call went to the row of its (variant, barcode): row products are in (0, 1]
    assert bool(((pack.csc_e >= 0) & (pack.csc_e <= 1)).all())
    # schedule: a permutation of the barcodes by non-increasing depth
    depth = (bo[1:] - bo[:-1])[pack.barcode_order.to(torch.int64)]
    assert bool((depth[1:] <= depth[:-1]).all())
    assert bool((torch.sort(pack.barcode_order).values == torch.arange(pack.n_barcodes, device=pack.device)).all())


def test_first_barcodes_match_the_oracle(full):
    from bench import slice_barcodes
    ds, pack, D = full
    n_b = 24
    table = D._probs_table(pack, None, 0.01)
    logits, post, _ = D._e_step(pack, table, 0.35)
    got_l, got_p = logits[:n_b].cpu().numpy(), post[:n_b].cpu().numpy()
    calls, handler = slice_barcodes(ds, n_b)
    # same regularised betas as the full run (learn mode counts molecules of every barcode): reuse the table
    _, _, _, rows = oracle.OracleDemultiplexer.pack_calls(calls, ds.genotypes, False)
    table_host = table.cpu().numpy()[:, :ds.genotypes.n_genotypes]
    want_l = oracle.barcode_logits(rows['variant_id'], rows['compressed_cb'], rows['p_base_wrong'], table_host, 0.35, n_b)
    want_p = oracle.softmax_rows(want_l)
    rel = np.abs(got_l - want_l) / np.maximum(np.abs(want_l), 1e-30)
    assert rel.max() <= 1e-5, rel.max()
    dlogit = np.abs(got_l.astype(np.float64) - want_l).max(axis=1, keepdims=True)
    assert (np.abs(got_p - want_p) <= 1e-6 + 0.5 * dlogit).all()
    assert (np.argmax(got_p, 1) == np.argmax(want_p, 1)).all()
    assert np.allclose(post.sum(dim=1).cpu().numpy(), 1, atol=1e-4)


def test_shards_and_mstep_additivity_at_full_size(full):
    import torch
    ds, pack, D = full
    B = pack.n_barcodes
    table = D._probs_table(pack, None, 0.01)
    logits, _, singlets = D._e_step(pack, table, 0.35, want_post=False, want_singlets=True)
    whole = D._m_step(pack, singlets)
    assert bool(torch.isfinite(whole).all()) and bool((whole >= 0).all())
    total64 = torch.zeros((pack.n_variants, pack.n_genotypes), dtype=torch.float64, device=pack.device)
    from demuxalot_b200 import _native
    lib = _native.load()
    for lo, hi in ((0, B // 2), (B // 2, B)):
        shard = D._pack_device(ds.calls, ds.genotypes, B, add_data_prior=True, barcode_range=(lo, hi))
        assert shard.n_barcodes == hi - lo and shard.barcode_range == (lo, hi)  # barcode ids local to the range
        shard_logits, _, _ = D._e_step(shard, table, 0.35, want_post=False)
        assert torch.equal(shard_logits, logits[lo:hi])  # the E-step is barcode-local: bit-identical
        part64 = torch.empty_like(total64)
        assert lib.dmx_mstep(shard.variant_offsets.data_ptr(), shard.csc_cb.data_ptr(), shard.csc_e.data_ptr(),
                             singlets[lo:hi].data_ptr(), singlets.shape[1], pack.n_genotypes, 2.0, 0, 0, part64.data_ptr(),
                             pack.n_genotypes, 0, pack.n_variants, torch.cuda.current_stream().cuda_stream) == 0
        total64 += part64
        del shard
    summed = total64.to(torch.float32)
    assert float((summed != whole).float().mean()) < 1e-4  # float64 regrouping flips a float32 bit only rarely
    assert torch.allclose(summed, whole, rtol=1e-6, atol=0)


def test_learn_genotypes_runs_at_full_size(full):
    ds, pack, D = full
    learnt, post = D.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=3, doublet_prior=0.35)
    betas = np.array(learnt.get_betas())
    raw = np.array(ds.genotypes.get_betas())
    assert betas.shape == raw.shape and np.isfinite(betas).all() and (betas >= raw).all()
    assert post.shape == (10_000, 528) and np.allclose(post.values.sum(axis=1), 1, atol=1e-4)
    # donors are recovered: singlet barcodes with decent depth are called correctly
    donors = ds.barcode_donors
    singlet = donors[:, 1] < 0
    called = post.values[:, :32].argmax(axis=1)
    confident = singlet & (post.values[:, :32].max(axis=1) > 0.9)
    assert confident.sum() > 0.5 * singlet.sum()
    assert (called[confident] == donors[confident, 0]).mean() > 0.99
