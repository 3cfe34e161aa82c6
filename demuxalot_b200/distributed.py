"""
Multi-GPU EM: one process per GPU (torch.distributed, NCCL over NVLink), work partitioned by barcode.

The E-step is barcode-local (a barcode's rows and the replicated probability table are all it reads), so it
needs no communication.  The M-step on a barcode shard yields a partial variant x genotype sum; the partials are
combined with ONE all-reduce per EM iteration, pipelined against the M-step kernel over
tiles of the variant range (`Demultiplexer._m_step`).  Partials travel as float64 and are rounded to float32 once
after the global sum, exactly where the single-GPU path rounds.  The only other exchange is a one-off integer
all-reduce of the per-variant molecule counts that enter the data prior (demux.py:381).

Two ways to use it (both need `torch.distributed` initialised, one rank per GPU):

  * sharded(): every rank is given the SAME full inputs; rank r keeps the calls of its contiguous barcode range
    (balanced by call count) and returns the posteriors of that range, optionally gathered to all ranks;
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


def learn_genotypes_sharded(chromosome2compressed_snp_calls, genotypes, barcode_handler, n_iterations=5,
                            p_genotype_clip=0.01, doublet_prior=0., barcode_prior_logits: np.ndarray = None,
                            group=None, gather_posteriors: bool = True):
    """
    Barcode-sharded `learn_genotypes`: same arguments on every rank, same learnt genotypes on every rank.
    Returns (learnt genotypes, posteriors DataFrame).  With gather_posteriors the frame covers all barcodes on
    every rank; otherwise only the rows of this rank's barcode range.
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

    shards = plan_barcode_shards(calls_per_barcode(chromosome2compressed_snp_calls, n_barcodes), world)
    lo, hi = shards[rank]
    with em_group(group):
        # every rank sees every call, so the molecule counts of the data prior are already global: no all-reduce
        saved, Demultiplexer.process_group = Demultiplexer.process_group, None
        try:
            pack = Demultiplexer._pack_device(chromosome2compressed_snp_calls, genotypes, n_barcodes,
                                              add_data_prior=True, barcode_range=(lo, hi))
        finally:
            Demultiplexer.process_group = saved
        pack.check()
        prior_dev = Demultiplexer._prior_logits_to_device(barcode_prior_logits, n_barcodes, n_cols, pack.device)
        post, addition = Demultiplexer._em_iterations(pack, n_iterations, p_genotype_clip, doublet_prior, prior_dev)
    names = option_names(genotypes.genotype_names, doublet_prior)
    learnt = genotypes._with_betas((pack.raw_betas + addition).cpu().numpy())
    if gather_posteriors:
        # barcodes outside a rank's range have no rows there; every rank contributes its own block
        counts = [h - l for l, h in shards]
        blocks = [torch.empty((c, n_cols), dtype=torch.float32, device=pack.device) for c in counts]
        mine = post[lo:hi].contiguous()
        dist.all_gather(blocks, mine, group=group) if len(set(counts)) == 1 else _all_gather_ragged(blocks, mine, group)
        full = torch.cat(blocks, dim=0).cpu().numpy()
        return learnt, pd.DataFrame(data=full, index=barcode_handler.ordered_barcodes, columns=names)
    return learnt, pd.DataFrame(data=post[lo:hi].cpu().numpy(), index=barcode_handler.ordered_barcodes[lo:hi],
                                columns=names)


def _all_gather_ragged(blocks, mine, group) -> None:
    """all_gather for blocks of different row counts: one broadcast per rank (blocks are small: [B/N, C])."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    for src, block in enumerate(blocks):
        if src == rank:
            block.copy_(mine)
        dist.broadcast(block, src=dist.get_global_rank(group, src), group=group)
