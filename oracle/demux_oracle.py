"""
TEST INFRASTRUCTURE -- CPU restatement (numpy) of demuxalot's likelihood / EM core.

This file is the *checker* for the CUDA path, never the thing that is shipped or measured as
the product (see oracle/__init__.py).  Every function cites the reference lines it restates
(paths relative to the upstream repository, v0.4.3).  The arithmetic (dtypes, order of
roundings, float64 accumulation followed by one rounding to float32) follows the reference
exactly; the data handling is re-expressed over plain struct-of-arrays inputs with 64-bit
integer keys instead of numpy structured arrays, so the same functions can be fed either the
reference's objects or ours (duck typing on attribute names).

Parity: pinned against fixtures produced by the unmodified reference
(tests/golden/make_golden.py -> tests/golden/*.npz) and, when /root/reference is present,
against the live reference (tests/test_oracle_vs_reference.py).
"""
from __future__ import annotations

import multiprocessing as mp
import os
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import pandas as pd

BASE_CODES = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'N': 4}  # demuxalot/utils.py:24

ERROR_FLOOR = 1e-4  # demux.py:261  `.clip(1e-4)`
DENOM_FLOOR = 1e-7  # demux.py:273  `denom.clip(1e-7)`
PRIOR_REGULARISATION = 100.  # demux.py:382-383
PRIOR_BASELINE = 1.  # demux.py:377


# ----------------------------------------------------------------------------------------------
# columns: singlets, then doublets                                   demux.py:158-191
# ----------------------------------------------------------------------------------------------

def n_options(n_genotypes: int, doublet_prior: float) -> int:
    return n_genotypes if doublet_prior == 0 else n_genotypes * (n_genotypes + 1) // 2


def doublet_penalties(n_genotypes: int, doublet_prior: float) -> np.ndarray:
    """float32 [C] additive logit offsets; restates Demultiplexer._doublet_penalties (demux.py:158-173)."""
    assert 0 <= doublet_prior < 1
    out = np.zeros(n_options(n_genotypes, doublet_prior), dtype=np.float32)
    if doublet_prior != 0:
        g = n_genotypes
        bonus = np.log(g * doublet_prior)
        bonus = bonus - np.log(g * max(g - 1, 1) / 2 * (1 - doublet_prior))
        out[g:] = bonus  # float64 -> float32 rounding happens here, as in the reference
    return out


def option_pairs(n_genotypes: int, doublet_prior: float) -> np.ndarray:
    """int32 [C, 2]: (i, i) for singlets, then (i, j), i < j, i-major; order of demux.py:179-191."""
    g = n_genotypes
    pairs = [(i, i) for i in range(g)]
    if doublet_prior != 0:
        pairs += [(i, j) for i in range(g) for j in range(i + 1, g)]
    return np.asarray(pairs, dtype=np.int32).reshape(-1, 2)


def option_names(genotype_names: Iterable[str], doublet_prior: float) -> List[str]:
    names = list(genotype_names)
    out = list(names)
    if doublet_prior != 0:
        out += [f'{a}+{b}' for ia, a in enumerate(names) for ib, b in enumerate(names) if ia < ib]
    return out


# ----------------------------------------------------------------------------------------------
# ids                                                                genotypes.py:56-66
# ----------------------------------------------------------------------------------------------

def snp_ids_for_variants(var2varid: Dict[Tuple, int]) -> np.ndarray:
    """variant row -> dense id of its (chrom, pos), ids in first-seen order over the dict."""
    snp2id: Dict[Tuple, int] = {}
    out = np.full(len(var2varid), -1, dtype=np.int32)
    for (chrom, pos, _base), vid in var2varid.items():
        out[vid] = snp2id.setdefault((chrom, pos), len(snp2id))
    assert np.all(out >= 0)
    return out


def _genotype_keys(var2varid: Dict[Tuple, int]):
    """64-bit key (chrom_id << 40 | pos << 8 | base) per variant + chromosome name -> id map."""
    chrom2id: Dict[object, int] = {}
    n = len(var2varid)
    keys = np.empty(n, dtype=np.int64)
    vids = np.empty(n, dtype=np.int32)
    for k, ((chrom, pos, base), vid) in enumerate(var2varid.items()):
        cid = chrom2id.setdefault(chrom, len(chrom2id))
        keys[k] = (cid << 40) | ((int(pos) & 0xFFFFFFFF) << 8) | BASE_CODES[base]
        vids[k] = vid
    # demux.py:317 -- variant indices must enumerate 0..V-1
    assert np.array_equal(np.sort(vids), np.arange(n)), 'variant ids are not a permutation of 0..V-1'
    order = np.argsort(keys, kind='stable')
    return keys[order], vids[order], chrom2id


def match_and_flatten_calls(chromosome2compressed_snp_calls, var2varid, variant2snp: np.ndarray):
    """
    Restates the matching part of Demultiplexer.pack_calls (demux.py:308-363).

    Returns the *matched* molecule-level calls, in dict order over chromosomes then call order:
    dict(variant_id i4, snp_id i4, compressed_cb i4, molecule_id i4, p_base_wrong f4,
         p_molecule_aligned_wrong f4).
    """
    gkeys, gvids, chrom2id = _genotype_keys(var2varid)
    parts = {k: [] for k in ('variant_id', 'compressed_cb', 'molecule_id', 'p_base_wrong',
                             'p_molecule_aligned_wrong')}
    for chrom, calls in chromosome2compressed_snp_calls.items():
        snp_calls = calls.snp_calls[:calls.n_snp_calls]
        molecules = calls.molecules[:calls.n_molecules]
        if chrom not in chrom2id:
            # demux.py:339-341 skips the chromosome, then the counter check at :359 fails
            assert calls.n_snp_calls == 0, f'calls on chromosome {chrom!r} that genotypes do not know'
            continue
        ckeys = (np.int64(chrom2id[chrom]) << 40) \
            | ((snp_calls['snp_position'].astype(np.int64) & 0xFFFFFFFF) << 8) \
            | snp_calls['base_index'].astype(np.int64)
        if len(gkeys):
            slot = np.searchsorted(gkeys, ckeys).clip(0, len(gkeys) - 1)
            vid = np.where(gkeys[slot] == ckeys, gvids[slot], -1).astype(np.int32)
        else:
            vid = np.full(len(ckeys), -1, dtype=np.int32)
        mol = snp_calls['molecule_index']
        parts['variant_id'].append(vid)
        parts['compressed_cb'].append(molecules['compressed_cb'][mol])
        parts['molecule_id'].append(mol)
        parts['p_base_wrong'].append(snp_calls['p_base_wrong'])
        parts['p_molecule_aligned_wrong'].append(molecules['p_group_misaligned'][mol])

    dtypes = dict(variant_id=np.int32, compressed_cb=np.int32, molecule_id=np.int32,
                  p_base_wrong=np.float32, p_molecule_aligned_wrong=np.float32)
    flat = {k: (np.concatenate(v) if v else np.zeros(0)).astype(dtypes[k]) for k, v in parts.items()}
    keep = flat['variant_id'] != -1  # demux.py:362-363
    flat = {k: v[keep] for k, v in flat.items()}
    flat['snp_id'] = variant2snp[flat['variant_id']].astype(np.int32)
    return flat


def group_molecule_calls(variant_id, snp_id, compressed_cb, p_base_wrong):
    """
    Restates Demultiplexer.molecule_calls2barcode_calls (demux.py:276-300).

    Unique (variant_id, snp_id, compressed_cb) in ascending lexicographic order (snp_id is a function
    of variant_id, so a (variant_id << 32 | cb) key gives the same order), p_base_wrong = float32
    product over the group's molecules in original call order starting from 1.0f (np.multiply.at
    semantics, demux.py:282-283), barcode_variant_count = group size, barcode_snp_count = number of
    molecule calls sharing (snp_id, cb) (demux.py:285-288).
    """
    variant_id = np.asarray(variant_id, dtype=np.int32)
    compressed_cb = np.asarray(compressed_cb, dtype=np.int32)
    e = np.asarray(p_base_wrong, dtype=np.float32)
    n = len(variant_id)
    if n == 0:
        z = np.zeros(0, dtype=np.int32)
        return dict(variant_id=z, snp_id=z.copy(), compressed_cb=z.copy(), p_base_wrong=np.zeros(0, np.float32),
                    barcode_variant_count=np.zeros(0, np.int64), barcode_snp_count=np.zeros(0, np.float64))
    key = (variant_id.astype(np.int64) << 32) | compressed_cb.astype(np.int64)
    order = np.argsort(key, kind='stable')
    skey = key[order]
    head = np.ones(n, dtype=bool)
    head[1:] = skey[1:] != skey[:-1]
    starts = np.flatnonzero(head)
    counts = np.diff(np.append(starts, n)).astype(np.int64)
    e_sorted = e[order]
    prod = e_sorted[starts].copy()  # 1.0f * e == e exactly
    for k in range(1, int(counts.max())):
        sel = np.flatnonzero(counts > k)
        prod[sel] = prod[sel] * e_sorted[starts[sel] + k]  # float32 multiply, left to right
    rows_variant = variant_id[order][starts]
    rows_cb = compressed_cb[order][starts]
    rows_snp = np.asarray(snp_id, dtype=np.int32)[order][starts]
    key2 = (rows_snp.astype(np.int64) << 32) | rows_cb.astype(np.int64)
    _, inv = np.unique(key2, return_inverse=True)
    snp_count = np.bincount(inv, weights=counts)[inv]
    return dict(variant_id=rows_variant, snp_id=rows_snp, compressed_cb=rows_cb, p_base_wrong=prod,
                barcode_variant_count=counts, barcode_snp_count=snp_count)


# ----------------------------------------------------------------------------------------------
# betas -> regularised betas -> probability table           demux.py:367-390, 267-274
# ----------------------------------------------------------------------------------------------

def _normalise_over_snp(values, variant2snp, regularisation):
    per_snp = np.bincount(variant2snp, weights=values)[variant2snp]  # float64
    return values / (per_snp + regularisation)


def regularised_betas(raw_betas: np.ndarray, variant2snp: np.ndarray, default_prior: float,
                      molecule_variant_ids: Optional[np.ndarray]) -> np.ndarray:
    """
    Restates compute_prior_betas (demux.py:372-385).  `molecule_variant_ids` = variant ids of the
    matched molecule-level calls when add_data_prior (learn_genotypes), None for predict_posteriors.
    """
    raw_betas = np.asarray(raw_betas)
    assert raw_betas.dtype == np.float32
    assert np.all(raw_betas >= 0), 'bad genotypes provided, negative betas appeared'
    n_variants = raw_betas.shape[0]
    prior = PRIOR_BASELINE
    if molecule_variant_ids is not None:
        n_mol = np.bincount(molecule_variant_ids, minlength=n_variants)
        prior = prior + _normalise_over_snp(n_mol, variant2snp, PRIOR_REGULARISATION)
    prior = prior + _normalise_over_snp(raw_betas.sum(axis=1), variant2snp, PRIOR_REGULARISATION)
    addition = prior[:, np.newaxis] * default_prior
    return raw_betas + addition.astype(np.float32)


def probs_from_betas(variant2snp: np.ndarray, betas: np.ndarray, p_genotype_clip: float) -> np.ndarray:
    """Restates Demultiplexer._compute_probs_from_betas (demux.py:267-274): float32 [V, G]."""
    betas = np.asarray(betas, dtype=np.float32)
    probs = np.zeros(betas.shape, dtype=np.float32)
    for g in range(betas.shape[1]):
        denom = np.bincount(variant2snp, weights=betas[:, g])[variant2snp]
        probs[:, g] = betas[:, g] / denom.clip(DENOM_FLOOR)
    return probs.clip(p_genotype_clip, 1 - p_genotype_clip)


# ----------------------------------------------------------------------------------------------
# E-step                                                       demux.py:246-265, utils.py:35-36
# ----------------------------------------------------------------------------------------------

_SHARED = {}


def _logit_columns(col_lo: int, col_hi: int) -> np.ndarray:
    s = _SHARED
    pairs, table = s['pairs'], s['table']
    vid, cb, one_minus_e, e_floor = s['vid'], s['cb'], s['one_minus_e'], s['e_floor']
    out = np.empty((s['n_barcodes'], col_hi - col_lo), dtype=np.float32)
    for k, c in enumerate(range(col_lo, col_hi)):
        i, j = pairs[c]
        col = table[:, i] if i == j else (table[:, i] + table[:, j]) * np.float32(0.5)  # demux.py:180,190
        p = col[vid]
        log_pen = np.log(p * one_minus_e + e_floor)  # float32 throughout, demux.py:261
        acc = s['penalties'][c] + np.zeros(s['n_barcodes'], dtype=np.float32)
        # utils.py:36 -- float32 + float64 bincount, rounded to float32 once on assignment
        out[:, k] = acc + np.bincount(cb, weights=log_pen, minlength=s['n_barcodes'])
    return out


def barcode_logits(rows_variant, rows_cb, rows_e, table: np.ndarray, doublet_prior: float,
                   n_barcodes: int, n_jobs: int = 1) -> np.ndarray:
    """
    Restates compute_barcode_logits_using_barcode_calls (demux.py:246-265): float32 [B, C].
    n_jobs > 1 shards the *columns* over forked worker processes (same arithmetic per column; used
    only to give the CPU baseline all host cores).
    """
    g = table.shape[1]
    e = np.asarray(rows_e, dtype=np.float32)
    _SHARED.update(
        pairs=option_pairs(g, doublet_prior), penalties=doublet_penalties(g, doublet_prior),
        table=np.asarray(table, dtype=np.float32), vid=np.asarray(rows_variant), cb=np.asarray(rows_cb),
        one_minus_e=1 - e, e_floor=e.clip(ERROR_FLOOR), n_barcodes=int(n_barcodes),
    )
    n_cols = len(_SHARED['pairs'])
    try:
        if n_jobs <= 1 or n_cols < 2 * n_jobs:
            return _logit_columns(0, n_cols)
        bounds = np.linspace(0, n_cols, 4 * n_jobs + 1).astype(int)
        with mp.get_context('fork').Pool(n_jobs) as pool:
            chunks = pool.starmap(_logit_columns, [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a])
        return np.concatenate(chunks, axis=1)
    finally:
        _SHARED.clear()


def softmax_rows(logits: np.ndarray) -> np.ndarray:
    """scipy.special.softmax(x, axis=-1) as used at demux.py:101,152 (float32 in, float32 out)."""
    shifted = np.exp(logits - np.amax(logits, axis=-1, keepdims=True))
    return shifted / np.sum(shifted, axis=-1, keepdims=True)


# ----------------------------------------------------------------------------------------------
# per-(barcode, SNP) regularised likelihood: `aggregate_on_snps = True`   demux.py:204-244
# ----------------------------------------------------------------------------------------------

P_BAD_SNP = 0.01  # demux.py:234


def log_softmax_rows(x: np.ndarray) -> np.ndarray:
    """scipy.special.log_softmax(x, axis=1) (scipy 1.18.1, special/_logsumexp.py) in the dtype of x."""
    x_max = np.max(x, axis=1, keepdims=True)
    x_max[~np.isfinite(x_max)] = 0
    tmp = x - x_max
    with np.errstate(divide='ignore'):
        out = np.log(np.sum(np.exp(tmp), axis=1, keepdims=True))
    return tmp - out


def snp_groups(mol_cb, mol_snp):
    """
    Dense ids of the (compressed_cb, snp_id) groups, ascending in (barcode, SNP): what FeatureLookup
    (utils.py:207-265) yields at demux.py:216-218, over int64 keys.  Returns
    (group_of_call [M'], molecules per group [n_groups] int64, barcode of group [n_groups]).
    """
    cb, snp = np.asarray(mol_cb), np.asarray(mol_snp)
    n_snp_categories = int(np.max(snp)) + 1  # utils.py:213 (raises on zero matched calls, as the reference does)
    keys = cb.astype(np.int64) * n_snp_categories + snp
    lookup, group_of_call = np.unique(keys, return_inverse=True)
    counts = np.bincount(group_of_call, minlength=len(lookup))
    return group_of_call, counts, lookup // n_snp_categories


def snp_aggregated_logits(mol_variant, mol_snp, mol_cb, mol_e, table: np.ndarray, doublet_prior: float,
                          n_barcodes: int, compensation: float = 0.5) -> np.ndarray:
    """
    Restates the `aggregate_on_snps=True` branch of compute_barcode_logits (demux.py:204-244): float64 [B, C].
    Molecule-level calls are grouped per (barcode, SNP); a group's float32 log-likelihoods are divided by
    count**compensation, passed through log_softmax, mixed with a uniform `P_BAD_SNP` (float64 from here on,
    because np.log returns a float64 scalar), log_softmax-ed again and summed per barcode.  As in the
    reference, the doublet penalties only fix the number of columns (demux.py:212) and are never added.
    """
    g = table.shape[1]
    pairs = option_pairs(g, doublet_prior)
    n_cols = len(pairs)
    group_of_call, counts, group_barcode = snp_groups(mol_cb, mol_snp)
    n_groups = len(counts)
    table = np.asarray(table, dtype=np.float32)
    e = np.asarray(mol_e, dtype=np.float32)
    variant = np.asarray(mol_variant)
    group_logits = np.zeros((n_groups, n_cols), dtype=np.float32)
    for c, (i, j) in enumerate(pairs):
        col = table[:, i] if i == j else (table[:, i] + table[:, j]) * np.float32(0.5)  # demux.py:180,190
        log_pen = np.log(col[variant] + e)  # demux.py:227-228, float32
        group_logits[:, c] = group_logits[:, c] + np.bincount(group_of_call, weights=log_pen, minlength=n_groups)
    group_logits /= counts[:, None] ** compensation  # demux.py:231: float64 quotient stored as float32
    group_logits = log_softmax_rows(group_logits)  # demux.py:233, float32
    group_logits = np.logaddexp(group_logits, np.log(P_BAD_SNP / n_cols))  # demux.py:235 -> float64
    group_logits = log_softmax_rows(group_logits)  # demux.py:236, float64
    return np.stack([np.bincount(group_barcode, weights=col, minlength=n_barcodes) for col in group_logits.T],
                    axis=1)  # demux.py:238-241


# ----------------------------------------------------------------------------------------------
# M-step                                                             demux.py:113-118
# ----------------------------------------------------------------------------------------------

def m_step(rows_variant, rows_cb, rows_e, posteriors: np.ndarray, n_genotypes: int, n_variants: int,
           contribution_power: float = 2.) -> np.ndarray:
    """float32 [V, G] genotype_addition; only the singlet posterior columns are used (demux.py:115)."""
    addition = np.zeros((n_variants, n_genotypes), dtype=np.float32)
    one_minus_e = 1 - np.asarray(rows_e, dtype=np.float32)
    for g in range(n_genotypes):
        contribution = posteriors[rows_cb, g] * one_minus_e
        contribution **= contribution_power
        addition[:, g] = addition[:, g] + np.bincount(rows_variant, weights=contribution, minlength=n_variants)
    return addition


# ----------------------------------------------------------------------------------------------
# API-level restatement (same signatures / return objects as the reference's Demultiplexer)
# ----------------------------------------------------------------------------------------------

class OracleDemultiplexer:
    """Mirrors demuxalot.Demultiplexer (demux.py:24-392) on top of the functions above."""
    contribution_power = 2.
    aggregate_on_snps = False  # demux.py:31; True selects snp_aggregated_logits (demux.py:204-244)
    compensation_during_computing_barcode_logits = 0.5  # demux.py:32
    n_jobs = 1  # column-sharding of the E-step over processes (CPU baseline only)

    @staticmethod
    def pack_calls(chromosome2compressed_snp_calls, genotypes, add_data_prior: bool):
        variant2snp = snp_ids_for_variants(genotypes.var2varid)
        molecule_calls = match_and_flatten_calls(chromosome2compressed_snp_calls, genotypes.var2varid, variant2snp)
        rows = group_molecule_calls(molecule_calls['variant_id'], molecule_calls['snp_id'],
                                    molecule_calls['compressed_cb'], molecule_calls['p_base_wrong'])
        betas = regularised_betas(
            genotypes.get_betas(), variant2snp, genotypes.default_prior,
            molecule_calls['variant_id'] if add_data_prior else None)
        betas.flags.writeable = False
        return variant2snp, betas, molecule_calls, rows

    @classmethod
    def _logits(cls, rows, table, doublet_prior, n_barcodes, molecule_calls=None):
        if cls.aggregate_on_snps:  # demux.py:198
            return snp_aggregated_logits(
                molecule_calls['variant_id'], molecule_calls['snp_id'], molecule_calls['compressed_cb'],
                molecule_calls['p_base_wrong'], table, doublet_prior, n_barcodes,
                cls.compensation_during_computing_barcode_logits)
        return barcode_logits(rows['variant_id'], rows['compressed_cb'], rows['p_base_wrong'], table,
                              doublet_prior, n_barcodes, n_jobs=cls.n_jobs)

    @classmethod
    def predict_posteriors(cls, chromosome2compressed_snp_calls, genotypes, barcode_handler,
                           p_genotype_clip=0.01, doublet_prior=0.35):
        variant2snp, betas, _mol, rows = cls.pack_calls(chromosome2compressed_snp_calls, genotypes, False)
        table = probs_from_betas(variant2snp, betas, p_genotype_clip)
        assert np.isfinite(table).all()
        logits = cls._logits(rows, table, doublet_prior, barcode_handler.n_barcodes, _mol)
        names = option_names(genotypes.genotype_names, doublet_prior)
        index = list(barcode_handler.ordered_barcodes)
        logits_df = pd.DataFrame(data=logits, index=index, columns=names)
        logits_df.index.name = 'BARCODE'
        probs_df = pd.DataFrame(data=softmax_rows(logits), index=index, columns=names)
        probs_df.index.name = 'BARCODE'
        return logits_df, probs_df

    @classmethod
    def staged_genotype_learning(cls, chromosome2compressed_snp_calls, genotypes, barcode_handler,
                                 n_iterations=5, p_genotype_clip=0.01, doublet_prior=0.,
                                 barcode_prior_logits: np.ndarray = None):
        assert 0 <= doublet_prior < 1
        n_cols = n_options(genotypes.n_genotypes, doublet_prior)
        if barcode_prior_logits is not None:
            assert barcode_prior_logits.shape == (barcode_handler.n_barcodes, n_cols), 'wrong shape of priors'
        variant2snp, betas, _mol, rows = cls.pack_calls(chromosome2compressed_snp_calls, genotypes, True)
        names = option_names(genotypes.genotype_names, doublet_prior)
        addition = np.zeros_like(betas)
        for iteration in range(n_iterations):
            table = probs_from_betas(variant2snp, betas + addition, p_genotype_clip)
            logits = cls._logits(rows, table, doublet_prior, barcode_handler.n_barcodes, _mol)
            if iteration == 0 and barcode_prior_logits is not None:
                logits += barcode_prior_logits
            post = softmax_rows(logits)
            post_df = pd.DataFrame(data=post, index=barcode_handler.ordered_barcodes, columns=names)
            yield post_df, dict(barcode_logits=logits, genotype_prior=betas, genotype_addition=addition)
            addition = m_step(rows['variant_id'], rows['compressed_cb'], rows['p_base_wrong'], post,
                              genotypes.n_genotypes, betas.shape[0], cls.contribution_power)

    @classmethod
    def learn_genotypes(cls, chromosome2compressed_snp_calls, genotypes, barcode_handler, n_iterations=5,
                        p_genotype_clip=0.01, doublet_prior=0., barcode_prior_logits: np.ndarray = None):
        *_, (post_df, debug) = cls.staged_genotype_learning(
            chromosome2compressed_snp_calls, genotypes, barcode_handler, n_iterations=n_iterations,
            p_genotype_clip=p_genotype_clip, doublet_prior=doublet_prior,
            barcode_prior_logits=barcode_prior_logits)
        learnt = genotypes._with_betas(genotypes.get_betas() + debug['genotype_addition'])
        return learnt, post_df


def default_n_jobs() -> int:
    return max(1, len(os.sched_getaffinity(0)))
