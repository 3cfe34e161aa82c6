"""
N > 1 path.  On CPU (gloo, world_size 2) the host-side logic is covered: the shard plan, and the algebra of
barcode sharding -- per-shard M-step partials (computed here with the oracle's arithmetic), one float64 sum
all-reduce, one rounding -- against the unsharded result, plus the integer all-reduce of the data-prior molecule
counts in the one-lane-per-rank mode.  The NCCL version of the same check needs two GPUs (`-m gpu`).
"""
import os
import socket

import numpy as np
import pytest

from demuxalot_b200.distributed import calls_per_barcode, plan_barcode_shards


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_shard_plan_properties():
    rng = np.random.default_rng(0)
    for n, world in [(1000, 8), (10, 4), (3, 8), (0, 2), (57, 1)]:
        w = rng.lognormal(size=n)
        plan = plan_barcode_shards(w, world)
        assert len(plan) == world and plan[0][0] == 0 and plan[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(plan, plan[1:])) and all(lo <= hi for lo, hi in plan)
        if n >= 100:
            loads = np.array([w[lo:hi].sum() for lo, hi in plan])
            assert loads.max() <= w.sum() / world + w.max() + 1e-9  # off by at most one barcode
    assert plan_barcode_shards(np.ones(8), 4) == [(0, 2), (2, 4), (4, 6), (6, 8)]


def test_calls_per_barcode():
    from demuxalot_b200.synthetic import make_dataset
    ds = make_dataset(n_genotypes=3, n_snps=50, n_barcodes=20, rows_per_barcode=30, seed=5, spare_capacity=4)
    got = calls_per_barcode(ds.calls, 20)
    assert got.sum() == ds.n_calls and len(got) == 20


def _gloo_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    import torch
    import torch.distributed as dist
    import oracle
    from demuxalot_b200.synthetic import make_dataset
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ds = make_dataset(n_genotypes=6, n_snps=300, n_barcodes=64, rows_per_barcode=80, seed=31, shuffle_variants=True)
        O = oracle.OracleDemultiplexer
        v2s, betas, mol, rows = O.pack_calls(ds.calls, ds.genotypes, True)
        G, V, B = 6, betas.shape[0], 64
        table = oracle.probs_from_betas(v2s, betas, 0.01)
        logits = oracle.barcode_logits(rows['variant_id'], rows['compressed_cb'], rows['p_base_wrong'], table, 0.35, B)
        post = oracle.softmax_rows(logits)
        full = oracle.m_step(rows['variant_id'], rows['compressed_cb'], rows['p_base_wrong'], post, G, V)

        lo, hi = plan_barcode_shards(calls_per_barcode(ds.calls, B), world)[rank]
        mine = (rows['compressed_cb'] >= lo) & (rows['compressed_cb'] < hi)
        # the E-step is barcode-local: logits of my barcodes need only my rows
        my_logits = oracle.barcode_logits(rows['variant_id'][mine], rows['compressed_cb'][mine],
                                          rows['p_base_wrong'][mine], table, 0.35, B)
        assert np.array_equal(my_logits[lo:hi], logits[lo:hi])
        # float64 partial of the M-step over my rows (same terms as oracle.m_step, unrounded)
        w = 1 - rows['p_base_wrong'][mine]
        partial = np.zeros((V, G), dtype=np.float64)
        for g in range(G):
            contribution = post[rows['compressed_cb'][mine], g] * w
            contribution **= 2.
            partial[:, g] = np.bincount(rows['variant_id'][mine], weights=contribution, minlength=V)
        t = torch.from_numpy(partial)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        summed = t.numpy().astype(np.float32)
        assert np.allclose(summed, full, rtol=1e-6, atol=0)
        assert (summed == full).mean() > 0.999  # float64 regrouping may flip the last float32 bit, rarely

        # one lane per rank: the data prior needs the molecule counts of every lane (integer all-reduce)
        lane = make_dataset(n_genotypes=6, n_snps=300, n_barcodes=16, rows_per_barcode=40, seed=31, calls_seed=rank,
                            shuffle_variants=True)
        assert lane.genotypes.var2varid == ds.genotypes.var2varid  # same donors on every lane
        _, _, lane_mol, _ = O.pack_calls(lane.calls, lane.genotypes, True)
        counts = torch.from_numpy(np.bincount(lane_mol['variant_id'], minlength=V))
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        np.save(os.path.join(out_dir, f'counts_{rank}.npy'), counts.numpy())
        np.save(os.path.join(out_dir, f'lane_{rank}.npy'), np.bincount(lane_mol['variant_id'], minlength=V))
    finally:
        dist.destroy_process_group()


def test_barcode_sharding_algebra_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    c0, c1 = np.load(tmp_path / 'counts_0.npy'), np.load(tmp_path / 'counts_1.npy')
    assert np.array_equal(c0, c1)
    assert np.array_equal(c0, np.load(tmp_path / 'lane_0.npy') + np.load(tmp_path / 'lane_1.npy'))



def _route_numpy(variant, cb, e, cuts):
    """numpy statement of dmx_route_calls: matched calls grouped by owner rank, stable, barcode ids local."""
    keep = variant >= 0
    dest = np.searchsorted(np.asarray(cuts[1:-1]), cb, side='right')
    order = np.flatnonzero(keep)[np.argsort(dest[keep], kind='stable')]
    counts = np.bincount(dest[keep], minlength=len(cuts) - 1)
    return variant[order], cb[order] - np.asarray(cuts)[dest[order]], e[order], counts


def _exchange_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    import torch
    import torch.distributed as dist
    import oracle
    from demuxalot_b200.distributed import exchange_calls, plan_barcode_shards
    from demuxalot_b200.synthetic import make_dataset
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ds = make_dataset(n_genotypes=5, n_snps=200, n_barcodes=40, rows_per_barcode=60, seed=77)
        B = 40
        v2s = oracle.snp_ids_for_variants(ds.genotypes.var2varid)
        # every call in dict / call order with its variant (-1: unmatched), as dmx_unpack_match_calls leaves them
        variant, cb, e = [], [], []
        keys = {k: v for k, v in ds.genotypes.var2varid.items()}
        for chrom, c in ds.calls.items():
            sc, mols = c.snp_calls[:c.n_snp_calls], c.molecules[:c.n_molecules]
            lo, hi = c.n_snp_calls * rank // world, c.n_snp_calls * (rank + 1) // world  # this rank's slice
            for rec in sc[lo:hi]:
                variant.append(keys.get((chrom, int(rec['snp_position']), 'ACGTN'[rec['base_index']]), -1))
                cb.append(int(mols['compressed_cb'][rec['molecule_index']]))
                e.append(rec['p_base_wrong'])
        variant, cb, e = np.array(variant, np.int32), np.array(cb, np.int32), np.array(e, np.float32)
        hist = torch.from_numpy(np.bincount(cb[variant >= 0], minlength=B).astype(np.int64))
        dist.all_reduce(hist)
        ranges = plan_barcode_shards(hist.numpy(), world)
        cuts = [r[0] for r in ranges] + [B]
        sv, scb, se, counts = _route_numpy(variant, cb, e, cuts)
        received, n, bad = exchange_calls([torch.from_numpy(x.copy()) for x in (sv, scb, se)],
                                          [int(c) for c in counts], 0, dist.group.WORLD, 'cpu')
        assert bad == 0
        np.savez(os.path.join(out_dir, f'shard_{rank}.npz'), variant=received[0][:n].numpy(), cb=received[1][:n].numpy(),
                 e=received[2][:n].numpy(), lo=ranges[rank][0], hi=ranges[rank][1])
    finally:
        dist.destroy_process_group()


def test_sharded_pack_exchange_keeps_call_order_gloo_world2(tmp_path):
    """Slices of the calls -> route by barcode range -> all-to-all: every rank ends up with exactly the matched calls of
    its barcode range, call order kept inside every variant (what the ordered float32 products of demux.py:282-283
    need)."""
    import torch.multiprocessing as mp
    import oracle
    from demuxalot_b200.synthetic import make_dataset
    mp.spawn(_exchange_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ds = make_dataset(n_genotypes=5, n_snps=200, n_barcodes=40, rows_per_barcode=60, seed=77)
    v2s = oracle.snp_ids_for_variants(ds.genotypes.var2varid)
    flat = oracle.match_and_flatten_calls(ds.calls, ds.genotypes.var2varid, v2s)
    mv, mcb, me = flat['variant_id'], flat['compressed_cb'], flat['p_base_wrong']
    covered = 0
    for rank in range(2):
        got = np.load(tmp_path / f'shard_{rank}.npz')
        lo, hi = int(got['lo']), int(got['hi'])
        mine = (mcb >= lo) & (mcb < hi)
        # call order is kept inside every chromosome (slices of different chromosomes interleave by rank), hence
        # inside every variant and every (variant, barcode) group: compare after a stable sort by variant
        a, b = np.argsort(got['variant'], kind='stable'), np.argsort(mv[mine], kind='stable')
        assert np.array_equal(got['variant'][a], mv[mine][b])
        assert np.array_equal(got['cb'][a], (mcb[mine] - lo)[b])
        assert np.array_equal(got['e'][a].view(np.uint32), me[mine][b].view(np.uint32))
        covered += int(mine.sum())
    assert covered == len(mv)

def test_em_group_requires_initialised_process_group():
    from demuxalot_b200.distributed import em_group
    with pytest.raises(AssertionError):
        with em_group():
            pass


# ------------------------------------------------------------------------------------------------- two GPUs
def _nccl_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    import torch
    import torch.distributed as dist
    from demuxalot_b200 import Demultiplexer
    from demuxalot_b200.distributed import em_group, learn_genotypes_sharded
    from demuxalot_b200.synthetic import make_dataset
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        ds = make_dataset(n_genotypes=12, n_snps=1500, n_barcodes=200, rows_per_barcode=150, seed=41)
        learnt, post = learn_genotypes_sharded(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=4,
                                               doublet_prior=0.35)
        np.save(os.path.join(out_dir, f'betas_{rank}.npy'), np.array(learnt.get_betas()))
        np.save(os.path.join(out_dir, f'post_{rank}.npy'), post.values)
        # a different data set of the same table shape in between, then the first call again: the cached peer-mapped
        # tables are reused and must not leak one run's genotype_addition into the next one's first E-step
        other = make_dataset(n_genotypes=12, n_snps=1500, n_barcodes=90, rows_per_barcode=40, seed=41, calls_seed=7,
                             doublet_fraction=0.6)
        learn_genotypes_sharded(other.calls, other.genotypes, other.barcode_handler, n_iterations=3, doublet_prior=0.35)
        again, again_post = learn_genotypes_sharded(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=4,
                                                    doublet_prior=0.35)
        assert np.array_equal(np.array(again.get_betas()), np.array(learnt.get_betas())), 'stale state between EM runs'
        assert np.array_equal(again_post.values, post.values)
        if rank == 0:
            single, single_post = Demultiplexer.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler,
                                                                n_iterations=4, doublet_prior=0.35)
            np.save(os.path.join(out_dir, 'betas_single.npy'), np.array(single.get_betas()))
            np.save(os.path.join(out_dir, 'post_single.npy'), single_post.values)
        # lanes: each rank its own barcodes, shared donors
        lane = make_dataset(n_genotypes=12, n_snps=1500, n_barcodes=100, rows_per_barcode=150, seed=41, calls_seed=rank)
        with em_group():
            lane_learnt, _ = Demultiplexer.learn_genotypes(lane.calls, lane.genotypes, lane.barcode_handler,
                                                           n_iterations=3, doublet_prior=0.35)
        np.save(os.path.join(out_dir, f'lane_betas_{rank}.npy'), np.array(lane_learnt.get_betas()))
        # the sharded pack itself: this rank's rows = the oracle's rows of its barcode range, local barcode ids
        with em_group():
            pack = Demultiplexer._pack_device(ds.calls, ds.genotypes, 200, add_data_prior=True,
                                              shard=(rank, world, dist.group.WORLD))
        np.savez(os.path.join(out_dir, f'pack_{rank}.npz'), variant=pack.csc_variant.cpu().numpy(),
                 cb=pack.csc_cb.cpu().numpy(), e=pack.csc_e.cpu().numpy(), n_mol=pack.n_mol.cpu().numpy(),
                 betas=pack.betas.cpu().numpy(), lo=pack.barcode_range[0], hi=pack.barcode_range[1])
        # the NCCL exchange (dmx_mstep_allreduce), both wire formats, tiled; the runs above used the default exchange
        # (dmx_peer_sum_f32 over peer memory where the platform has symmetric memory)
        saved = (Demultiplexer.mstep_exchange, Demultiplexer.mstep_allreduce_dtype, Demultiplexer.mstep_allreduce_tiles)
        try:
            Demultiplexer.mstep_exchange = 'nccl'
            Demultiplexer.mstep_allreduce_dtype, Demultiplexer.mstep_allreduce_tiles = 'float32', 3
            narrow, _ = learn_genotypes_sharded(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=4,
                                                doublet_prior=0.35)
            Demultiplexer.mstep_allreduce_dtype, Demultiplexer.mstep_allreduce_tiles = 'float64', 2
            wide, _ = learn_genotypes_sharded(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=4,
                                              doublet_prior=0.35)
        finally:
            Demultiplexer.mstep_exchange, Demultiplexer.mstep_allreduce_dtype, Demultiplexer.mstep_allreduce_tiles = saved
        np.save(os.path.join(out_dir, f'betas_f32_{rank}.npy'), np.array(narrow.get_betas()))
        np.save(os.path.join(out_dir, f'betas_f64_{rank}.npy'), np.array(wide.get_betas()))
        # One M-step on the same posteriors under every exchange: the float64 wire must reproduce the single-GPU sums
        # (rounded once, after the global sum) while float32 partials visibly do not -- a switch that silently does
        # nothing fails here (it once did: every mode ran the float32 all-reduce and all parity bars still passed).
        with em_group():
            shard = Demultiplexer._pack_device(ds.calls, ds.genotypes, 200, add_data_prior=True,
                                               shard=(rank, world, dist.group.WORLD), keep_calls=False)
            table = Demultiplexer._probs_table(shard, None, 0.01)
            _, _, singlets = Demultiplexer._e_step(shard, table, 0.35, want_logits=False, want_post=False,
                                                   want_singlets=True)
            for exchange, wire in (('auto', 'float32'), ('nccl', 'float64'), ('nccl', 'float32')):
                saved = (Demultiplexer.mstep_exchange, Demultiplexer.mstep_allreduce_dtype)
                Demultiplexer.mstep_exchange, Demultiplexer.mstep_allreduce_dtype = exchange, wire
                try:
                    mbuf = Demultiplexer._mstep_buffers(shard)
                    summed = Demultiplexer._m_step(shard, singlets, out=mbuf['tables'][1], buffers=mbuf).clone()
                finally:
                    Demultiplexer.mstep_exchange, Demultiplexer.mstep_allreduce_dtype = saved
                tag = mbuf['last_exchange'].split('/tiles')[0].replace('/', '_')
                np.save(os.path.join(out_dir, f'mstep_{tag}_{rank}.npy'), summed.cpu().numpy())
        if rank == 0:
            full = Demultiplexer._pack_device(ds.calls, ds.genotypes, 200, add_data_prior=True, keep_calls=False)
            table = Demultiplexer._probs_table(full, None, 0.01)
            _, _, singlets = Demultiplexer._e_step(full, table, 0.35, want_logits=False, want_post=False,
                                                   want_singlets=True)
            np.save(os.path.join(out_dir, 'mstep_single.npy'), Demultiplexer._m_step(full, singlets).cpu().numpy())
    finally:
        from demuxalot_b200.distributed import release_native_comms
        release_native_comms()
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_sharded_em_matches_single_gpu(tmp_path, native_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    mp.spawn(_nccl_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    b0, b1, bs = (np.load(tmp_path / f) for f in ('betas_0.npy', 'betas_1.npy', 'betas_single.npy'))
    assert np.array_equal(b0, b1)
    assert np.allclose(b0, bs, rtol=1e-6, atol=1e-7) and (b0 == bs).mean() > 0.99
    p0, p1, ps = (np.load(tmp_path / f) for f in ('post_0.npy', 'post_1.npy', 'post_single.npy'))
    assert np.array_equal(p0, p1) and np.abs(p0 - ps).max() <= 1e-6
    assert np.array_equal(np.load(tmp_path / 'lane_betas_0.npy'), np.load(tmp_path / 'lane_betas_1.npy'))
    n0, n1 = np.load(tmp_path / 'betas_f32_0.npy'), np.load(tmp_path / 'betas_f32_1.npy')
    assert np.array_equal(n0, n1) and np.allclose(n0, bs, rtol=2e-6, atol=1e-7)
    w0, w1 = np.load(tmp_path / 'betas_f64_0.npy'), np.load(tmp_path / 'betas_f64_1.npy')
    assert np.array_equal(w0, w1) and np.allclose(w0, bs, rtol=1e-6, atol=1e-7) and (w0 == bs).mean() > 0.999
    # the exchanges on one M-step: every mode ran under its own name (the file names come from what ran), all ranks hold
    # the same bits, and the wire formats are distinguishable
    single_sums = np.load(tmp_path / 'mstep_single.npy')
    touched = single_sums != 0
    mismatch = {}
    for tag in ('peer_float32', 'nccl_float64', 'nccl_float32'):
        a, b = np.load(tmp_path / f'mstep_{tag}_0.npy'), np.load(tmp_path / f'mstep_{tag}_1.npy')
        assert np.array_equal(a, b), tag
        assert np.allclose(a, single_sums, rtol=1e-6, atol=0), tag
        mismatch[tag] = float((a[touched] != single_sums[touched]).mean())
    assert mismatch['nccl_float64'] < 1e-4, mismatch  # one rounding after the global sum: the bits of one GPU
    assert mismatch['nccl_float32'] > 10 * mismatch['nccl_float64'] + 1e-3, mismatch  # a rounding per shard shows
    assert mismatch['peer_float32'] > 10 * mismatch['nccl_float64'] + 1e-3, mismatch
    # against the oracle: learnt betas, the shards' rows and the data prior
    import oracle
    from demuxalot_b200.synthetic import make_dataset
    ds = make_dataset(n_genotypes=12, n_snps=1500, n_barcodes=200, rows_per_barcode=150, seed=41)
    O = oracle.OracleDemultiplexer
    want, _ = O.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=4, doublet_prior=0.35)
    wb = np.array(want.get_betas(), np.float64)
    assert (np.abs(b0 - wb) / np.maximum(np.abs(wb), 1e-3)).max() <= 1e-5
    _, obetas, omol, orows = O.pack_calls(ds.calls, ds.genotypes, True)
    n_rows = 0
    for rank in range(2):
        got = np.load(tmp_path / f'pack_{rank}.npz')
        lo, hi = int(got['lo']), int(got['hi'])
        mine = (orows['compressed_cb'] >= lo) & (orows['compressed_cb'] < hi)
        assert np.array_equal(got['variant'], orows['variant_id'][mine])
        assert np.array_equal(got['cb'], orows['compressed_cb'][mine] - lo)
        assert np.array_equal(got['e'].view(np.uint32), orows['p_base_wrong'][mine].view(np.uint32))
        assert np.array_equal(got['n_mol'], np.bincount(omol['variant_id'], minlength=len(obetas)))  # summed over ranks
        assert np.array_equal(got['betas'].view(np.uint32), obetas.view(np.uint32))
        n_rows += int(mine.sum())
    assert n_rows == len(orows['variant_id'])
    # lanes: the reference's equivalent is one run over the union of the lanes (BarcodeHandler over all of them)
    from demuxalot_b200 import CompressedSNPCalls
    lanes = [make_dataset(n_genotypes=12, n_snps=1500, n_barcodes=100, rows_per_barcode=150, seed=41, calls_seed=r)
             for r in range(2)]
    union = {}
    for chrom in lanes[0].calls:
        parts = []
        for k, lane in enumerate(lanes):
            c = lane.calls[chrom]
            shifted = CompressedSNPCalls.__new__(CompressedSNPCalls)
            shifted.molecules = c.molecules[:c.n_molecules].copy()
            shifted.molecules['compressed_cb'] += 100 * k
            shifted.snp_calls, shifted.n_molecules, shifted.n_snp_calls = c.snp_calls[:c.n_snp_calls].copy(), c.n_molecules, c.n_snp_calls
            parts.append(shifted)
        union[chrom] = CompressedSNPCalls.concatenate(parts)

    class UnionHandler:
        n_barcodes = 200
        ordered_barcodes = [str(k) for k in range(200)]

    want_lanes, _ = O.learn_genotypes(union, lanes[0].genotypes, UnionHandler, n_iterations=3, doublet_prior=0.35)
    wl = np.array(want_lanes.get_betas(), np.float64)
    lb = np.load(tmp_path / 'lane_betas_0.npy')
    assert (np.abs(lb - wl) / np.maximum(np.abs(wl), 1e-3)).max() <= 1e-5
