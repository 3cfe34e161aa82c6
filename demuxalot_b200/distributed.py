"""
Multi-GPU EM: one process per GPU (torch.distributed, NCCL over NVLink), work partitioned by barcode.

The E-step is barcode-local (a barcode's rows and the replicated probability table are all it reads), so it
needs no communication.  The M-step on a barcode shard yields a partial variant x genotype sum; the partials are
combined ONCE per EM iteration by `dmx_mstep_allreduce` (csrc/comm.cu): the M-step kernel
runs over tiles of the variant range and NCCL sums tile k while the kernel computes tile k + 1.  Partials travel as
float64 (reduce-scatter), are rounded to float32 once after the global sum, exactly where the single-GPU path
rounds, and the float32 slices are all-gathered.  The only other exchange is a one-off integer
all-reduce of the per-variant molecule counts that enter the data prior (demux.py:381).

Two ways to use it (both need `torch.distributed` initialised, one rank per GPU):

  * learn_genotypes_sharded(): every rank is given the SAME full inputs; rank r uploads and matches the r-th slice
    of the calls, the matched calls are exchanged over NVLink (one all-to-all) so that each rank sorts and keeps one
    contiguous barcode range (balanced by matched calls), and the posteriors of that range are returned, optionally
    gathered to all ranks;
  * lanes: every rank is given ITS OWN calls / barcode handler (e.g. one 10x lane per GPU sharing the donors);
    wrap the usual `Demultiplexer.learn_genotypes` call in `with em_group(group):`.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import List, Optional, Tuple

import numpy as np


def plan_barcode_shards(weight_per_barcode: np.ndarray, world_size: int) -> List[Tuple[int, int]]:
    """
    Contiguous barcode ranges [lo, hi) with near-equal total weight (calls or rows per barcode).
    Deterministic, covers [0, B) exactly, ranges may be empty when B < world_size.
    """
    w = np.asarray(weight_per_barcode, dtype=np.float64)
    n = len(w)
    if world_size <= 0:
        raise ValueError('world_size must be positive')
    csum = np.concatenate([[0.0], np.cumsum(w)])
    total = csum[-1]
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        cut = int(np.searchsorted(csum, target, side='left'))
        cuts.append(min(max(cut, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def calls_per_barcode(chromosome2compressed_snp_calls, n_barcodes: int) -> np.ndarray:
    """Molecule-level calls per barcode (host, one bincount per chromosome): the weight used for sharding."""
    out = np.zeros(n_barcodes, dtype=np.int64)
    for calls in chromosome2compressed_snp_calls.values():
        cb = calls.molecules['compressed_cb'][:calls.n_molecules]
        mol = calls.snp_calls['molecule_index'][:calls.n_snp_calls]
        if len(mol):
            out += np.bincount(cb[mol], minlength=n_barcodes)[:n_barcodes]
    return out


@contextmanager
def em_group(group=None):
    """Run `Demultiplexer.learn_genotypes` / `staged_genotype_learning` as one rank of a multi-GPU EM."""
    import torch.distributed as dist
    from .demultiplexer import Demultiplexer
    assert dist.is_initialized(), 'torch.distributed must be initialised (one process per GPU)'
    previous = Demultiplexer.process_group
    Demultiplexer.process_group = dist.group.WORLD if group is None else group
    try:
        yield
    finally:
        Demultiplexer.process_group = previous


def exchange_calls(send, counts: List[int], n_bad: int, group, device):
    """
    All-to-all of the routed calls (sharded pack).  `send`: tensors whose first sum(counts) entries are grouped by
    destination rank (counts[d] entries for rank d).  Returns (received tensors, n_received, total n_bad): blocks
    arrive in source-rank order and ranks hold consecutive slices of every chromosome's calls, so call order is kept
    inside every chromosome -- hence inside every (variant, barcode) group, which is all demux.py:282-283 needs.  Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests).
    """
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    meta = torch.tensor(list(counts) + [n_bad], dtype=torch.int64, device=device)
    gathered = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(gathered, meta, group=group)
    table = torch.stack(gathered).cpu()
    total_bad = int(table[:, world].sum())
    recv_counts = [int(table[src, rank]) for src in range(world)]
    n_recv, n_send = sum(recv_counts), sum(counts)
    received = []
    for t in send:
        r = torch.empty(max(n_recv, 1), dtype=t.dtype, device=device)
        if total_bad == 0:
            dist.all_to_all_single(r[:n_recv], t[:n_send], recv_counts, list(counts), group=group)
        received.append(r)
    return received, n_recv, total_bad


_NATIVE_COMMS: dict = {}


def native_comm(group, device):
    """The library's own NCCL communicator for `group` (dmx_comm_init), created on first use; None when the group
    does not run on NCCL (then `Demultiplexer._m_step` goes through torch.distributed)."""
    import ctypes as C

    import torch
    import torch.distributed as dist
    from . import _native
    if dist.get_backend(group) != 'nccl':
        return None
    key = (id(group), torch.device(device).index)
    if key not in _NATIVE_COMMS:
        lib = _native.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        unique_id = (C.c_uint8 * 128)()
        if rank == 0:
            _native.check(lib.dmx_comm_unique_id(unique_id), 'dmx_comm_unique_id')
        box = [bytes(unique_id)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0), group=group)
        unique_id = (C.c_uint8 * 128).from_buffer_copy(box[0])
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _native.check(lib.dmx_comm_init(unique_id, rank, world, C.byref(handle)), 'dmx_comm_init')
        _NATIVE_COMMS[key] = handle
    return _NATIVE_COMMS[key]


_PEER_TABLES: dict = {}


def peer_tables(group, device, n_rows: int, n_cols: int):
    """
    Three float32 [n_rows, n_cols] tables that every rank of `group` has mapped into its address space (torch
    symmetric memory: CUDA IPC / fabric handles, i.e. NVLink peer memory) -- the M-step partial and the two alternating
    `genotype_addition` tables that `dmx_peer_sum_f32` reads from / writes to on all ranks.  Collective (every rank must
    call it with the same shape); cached per shape; None when the platform has no symmetric memory (agreed on by all
    ranks, `Demultiplexer._m_step` then sums through NCCL).
    """
    import torch
    import torch.distributed as dist
    key = (id(group), torch.device(device).index, n_rows, n_cols)
    if key in _PEER_TABLES:
        return _PEER_TABLES[key]
    entry, ok = None, 1
    try:
        import torch.distributed._symmetric_memory as symm_mem
        tensors = [symm_mem.empty((n_rows, n_cols), dtype=torch.float32, device=device) for _ in range(3)]
        handles = [symm_mem.rendezvous(t, group) for t in tensors]
        for t in tensors:
            t.zero_()
        entry = dict(partial=tensors[0], tables=tensors[1:], handles=handles, rank=dist.get_rank(group),
                     world=dist.get_world_size(group),
                     pointers={t.data_ptr(): [int(x) for x in h.buffer_ptrs] for t, h in zip(tensors, handles)})
    except Exception:  # noqa: BLE001 -- no symmetric memory on this platform / build
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        entry = None
    _PEER_TABLES[key] = entry
    return entry


def release_peer_tables() -> None:
    """Drops the cached peer-mapped tables (they are as large as the genotype table: 3 x V x G x 4 bytes)."""
    _PEER_TABLES.clear()


def release_native_comms() -> None:
    _PEER_TABLES.clear()
    from . import _native
    lib = _native.load()
    for handle in _NATIVE_COMMS.values():
        lib.dmx_comm_destroy(handle)
    _NATIVE_COMMS.clear()


def learn_genotypes_sharded(chromosome2compressed_snp_calls, genotypes, barcode_handler, n_iterations=5,
                            p_genotype_clip=0.01, doublet_prior=0., barcode_prior_logits: np.ndarray = None,
                            group=None, gather_posteriors: bool = True, device_parts=None):
    """
    Barcode-sharded `learn_genotypes`: same arguments on every rank, same learnt genotypes on every rank.
    Rank r uploads and matches the r-th slice of every chromosome's calls; the matched calls are exchanged so that
    each rank sorts and keeps one contiguous barcode range (balanced by matched calls); per EM iteration the M-step
    partials are summed across the ranks (`Demultiplexer._m_step`).
    Returns (learnt genotypes, posteriors DataFrame).  With gather_posteriors the frame covers all barcodes on
    every rank; otherwise only the rows of this rank's barcode range.
    `device_parts`: this rank's share of the calls already on its GPU (see Demultiplexer._unpack_device_parts)
    instead of `chromosome2compressed_snp_calls`.
    """
    import pandas as pd
    import torch
    import torch.distributed as dist
    from .demultiplexer import Demultiplexer, n_options, option_names

    assert dist.is_initialized(), 'torch.distributed must be initialised (one process per GPU)'
    group = dist.group.WORLD if group is None else group
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_barcodes = barcode_handler.n_barcodes
    n_cols = n_options(genotypes.n_genotypes, doublet_prior)
    assert 0 <= doublet_prior < 1 and n_iterations >= 1
    if barcode_prior_logits is not None:
        assert barcode_prior_logits.shape == (n_barcodes, n_cols), 'wrong shape of priors'

    with em_group(group):
        pack = Demultiplexer._pack_device(chromosome2compressed_snp_calls, genotypes, n_barcodes, add_data_prior=True,
                                          shard=(rank, world, group), device_parts=device_parts, keep_calls=False)
        pack.check()
        lo, hi = pack.barcode_range
        prior_dev = Demultiplexer._prior_logits_to_device(
            None if barcode_prior_logits is None else barcode_prior_logits[lo:hi], hi - lo, n_cols, pack.device)
        post, addition = Demultiplexer._em_iterations(pack, n_iterations, p_genotype_clip, doublet_prior, prior_dev)
    names = option_names(genotypes.genotype_names, doublet_prior)
    learnt = genotypes._with_betas((pack.raw_betas + addition).cpu().numpy())
    if gather_posteriors:
        # every rank contributes the block of its own barcode range
        ranges = [None] * world
        dist.all_gather_object(ranges, (lo, hi), group=group)
        blocks = [torch.empty((h - l, n_cols), dtype=torch.float32, device=pack.device) for l, h in ranges]
        _all_gather_ragged(blocks, post.contiguous(), group)
        full = torch.cat(blocks, dim=0).cpu().numpy()
        return learnt, pd.DataFrame(data=full, index=barcode_handler.ordered_barcodes, columns=names)
    return learnt, pd.DataFrame(data=post.cpu().numpy(), index=barcode_handler.ordered_barcodes[lo:hi], columns=names)


def _all_gather_ragged(blocks, mine, group) -> None:
    """all_gather for blocks of different row counts: one broadcast per rank (blocks are small: [B/N, C])."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    for src, block in enumerate(blocks):
        if src == rank:
            block.copy_(mine)
        dist.broadcast(block, src=dist.get_global_rank(group, src), group=group)
