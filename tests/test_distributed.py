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
        # float32 partial sums on the wire: one extra rounding per shard, otherwise the same protocol
        Demultiplexer.mstep_allreduce_dtype, Demultiplexer.mstep_allreduce_tiles = 'float32', 3
        try:
            narrow, _ = learn_genotypes_sharded(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=4,
                                                doublet_prior=0.35)
        finally:
            Demultiplexer.mstep_allreduce_dtype, Demultiplexer.mstep_allreduce_tiles = 'float64', 1
        np.save(os.path.join(out_dir, f'betas_f32_{rank}.npy'), np.array(narrow.get_betas()))
    finally:
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
    assert np.allclose(b0, bs, rtol=1e-6, atol=1e-7) and (b0 == bs).mean() > 0.999
    p0, p1, ps = (np.load(tmp_path / f) for f in ('post_0.npy', 'post_1.npy', 'post_single.npy'))
    assert np.array_equal(p0, p1) and np.abs(p0 - ps).max() <= 1e-6
    assert np.array_equal(np.load(tmp_path / 'lane_betas_0.npy'), np.load(tmp_path / 'lane_betas_1.npy'))
    n0, n1 = np.load(tmp_path / 'betas_f32_0.npy'), np.load(tmp_path / 'betas_f32_1.npy')
    assert np.array_equal(n0, n1) and np.allclose(n0, bs, rtol=2e-6, atol=1e-7)
