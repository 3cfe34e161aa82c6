"""
`detect_snps_positions` -- the documented "complex scenario" wrapper around the hot path (demuxalot/snp_detection.py:128-215,
SURVEY.md section 8(f) rank 2): a rough demultiplexing with the known genotypes (`predict_posteriors`, doublet prior 0 --
the singlet-only E-step kernel), then a scan of the BAM for positions where the donors assigned that way disagree.

Host-side glue, restated without pysam on top of `demuxalot_b200.bam` / `demuxalot_b200.counting`; the likelihood
call is the CUDA path.  Arithmetic, selection order and returned objects follow the reference line by line (cited
below), including its quirk that candidate positions of a fragment are indices relative to the fragment start
(snp_detection.py:51 -- only exact for fragments starting at 0, i.e. contigs shorter than `max_fragment_step`).
"""
from __future__ import annotations

from collections import Counter, defaultdict
from pathlib import Path
from typing import Dict, List

import numpy as np
import pandas as pd

from .bam import BamFile
from .barcodes import BarcodeHandler
from .calls import CompressedSNPCalls
from .counting import _bam_info, _open_cached, count_coverage_native, count_snps, parse_read as cellranger_parse_read
from .genotype_store import ProbabilisticGenotypes

_BASE_CODE = np.full(256, -1, dtype=np.int8)
for _k, _b in enumerate('ACGT'):
    _BASE_CODE[ord(_b)] = _k


def count_coverage(bamfile: BamFile, chromosome: str, start: int, stop: int, read_callback,
                   quality_threshold: int = 15) -> np.ndarray:
    """
    int32 [4, stop - start]: reads of A / C / G / T per reference position, what
    `np.asarray(pysam.AlignmentFile.count_coverage(chromosome, start=start, stop=stop, read_callback=...))` returns
    (snp_detection.py:37-42): reads overlapping the region that pass `read_callback`, aligned (M/=/X) bases only,
    base quality >= quality_threshold (pysam's default 15), N ignored.
    """
    cover = np.zeros((4, max(stop - start, 0)), dtype=np.int32)
    flat = cover.reshape(-1)
    width = cover.shape[1]
    for read in bamfile.fetch(chromosome, start, stop):
        if not read_callback(read):
            continue
        seq = read.seq
        if not seq:
            continue
        codes = _BASE_CODE[np.frombuffer(seq.encode(), dtype=np.uint8)]
        quals = np.frombuffer(bytes(read.query_qualities), dtype=np.uint8)
        qpos, rpos = 0, read.reference_start
        for op, length in read.cigartuples:
            if op in (0, 7, 8):
                lo, hi = max(rpos, start), min(rpos + length, stop)
                if hi > lo:
                    q0 = qpos + (lo - rpos)
                    base = codes[q0:q0 + (hi - lo)]
                    keep = base >= 0
                    if quality_threshold:
                        keep &= quals[q0:q0 + (hi - lo)] >= quality_threshold
                    index = base[keep].astype(np.int64) * width + (np.arange(lo, hi)[keep] - start)
                    np.add.at(flat, index, 1)  # a read covers a position once, but stay safe
                qpos += length
                rpos += length
            elif op in (1, 4):
                qpos += length
            elif op in (2, 3):
                rpos += length
            # hard clips / padding consume neither
    return cover


def _count_snp_stats_for_donors(compressed_snp_calls: CompressedSNPCalls, barcode_handler, barcode2donor, donor2dindex,
                                max_contribution_to_base_count_from_barcode=3.):
    """
    snp_detection.py:110-125: per position an int32 [n_donors, 4] matrix of base counts, every (barcode, position,
    base) contributing at most 3; confident calls only (p_base_wrong < 0.01).  Same dict order as the reference
    (first appearance of a position among the counted keys, keys in order of first appearance among the calls).
    """
    calls = compressed_snp_calls.snp_calls[:compressed_snp_calls.n_snp_calls]
    calls = calls[calls['p_base_wrong'] < 0.01]
    result: Dict[int, np.ndarray] = defaultdict(lambda: np.zeros([len(donor2dindex), 4], dtype='int32'))
    if len(calls) == 0:
        return result
    cb = compressed_snp_calls.molecules['compressed_cb'][calls['molecule_index']].astype(np.int64)
    pos = calls['snp_position'].astype(np.int64)
    position_type = calls['snp_position'].dtype.type  # the reference's dict keys are numpy int32 scalars
    base = calls['base_index'].astype(np.int64)
    keys = np.stack([cb, pos, base], axis=1)
    unique, first, counts = np.unique(keys, axis=0, return_index=True, return_counts=True)
    donor_of_cb = np.full(len(barcode_handler.ordered_barcodes), -1, dtype=np.int64)
    for barcode, donor in barcode2donor.items():
        donor_of_cb[barcode_handler.barcode2index[barcode]] = donor2dindex[donor]
    for k in np.argsort(first, kind='stable'):  # Counter order = order of first appearance
        donor = donor_of_cb[unique[k, 0]]
        if donor < 0:
            continue
        contribution = min(max_contribution_to_base_count_from_barcode, int(counts[k]))
        result[position_type(unique[k, 1])][donor, unique[k, 2]] += contribution  # base 4 (N) raises, as upstream
    return result


def detect_snps_for_chromosome(bamfile_path, chromosome, start, stop, sorted_donors, barcode2donor: dict, parse_read,
                               barcode_handler: BarcodeHandler, regularization: float, minimum_coverage: int,
                               minimum_alternative_fraction: float, minimum_alternative_coverage: int,
                               max_snp_candidates: int = 10000, minimum_fraction_of_ref_and_alt=0.98):
    """snp_detection.py:16-107: [(chromosome, position, importance per donor, {base: count})] for one fragment."""
    # stage 1: plain coverage, to find candidate positions
    coverage = 0
    bamfiles = [bamfile_path] if isinstance(bamfile_path, (str, Path)) else list(bamfile_path.values())
    for filename in bamfiles:
        native = count_coverage_native(filename, chromosome, start, stop, parse_read)  # built-in read filters only
        if native is None:
            native = count_coverage(_open_cached(filename), chromosome, start, stop,
                                    read_callback=lambda read: parse_read(read) is not None)
        coverage = coverage + native
    total = coverage.sum(axis=0)
    *_, alt, ref = np.sort(coverage, axis=0)
    is_candidate = (ref + alt) > minimum_coverage
    is_candidate &= (ref + alt) > minimum_fraction_of_ref_and_alt * total  # prefer SNPs with only two alternatives
    is_candidate &= alt > minimum_alternative_coverage
    is_candidate &= alt > ref * minimum_alternative_fraction
    candidate_positions = np.where(is_candidate)[0]  # relative to `start`, as in the reference (:51)
    if len(candidate_positions) > max_snp_candidates:
        candidate_positions = np.argsort(alt * is_candidate)[-max_snp_candidates:]
        candidate_positions = np.sort(candidate_positions)

    # stage 2: detailed counts at the candidates
    compressed_snp_calls = count_snps(bamfile_path, chromosome2positions={chromosome: candidate_positions},
                                      barcode_handler=barcode_handler, parse_read=parse_read, joblib_n_jobs=1,
                                      joblib_verbosity=0)
    if len(compressed_snp_calls) == 0:
        return []
    compressed_snp_calls = compressed_snp_calls[chromosome]
    donor2dindex = {donor: dindex for dindex, donor in enumerate(sorted_donors)}
    position2donor2base2count = _count_snp_stats_for_donors(compressed_snp_calls, barcode_handler, barcode2donor,
                                                            donor2dindex)

    def importance_and_base_counts(counts):  # counts: n_donors x 4 (snp_detection.py:79-100)
        top_bases = alt, ref = np.argsort(counts.sum(axis=0))[-2:]
        base_counts = {'ACGT'[ref]: counts[:, ref].sum(), 'ACGT'[alt]: counts[:, alt].sum()}
        counts = counts[:, top_bases] + 1e-4
        count_0, count_1 = counts.sum(axis=0)
        p_1_avg = count_1 / (count_1 + count_0)
        p_1 = (counts[:, 1] + p_1_avg * regularization) / (counts.sum(axis=1) + regularization)
        return np.square(p_1_avg - p_1), base_counts

    return [(chromosome, position) + importance_and_base_counts(counts)
            for position, counts in position2donor2base2count.items()]


def _select_top_snps(chrom_pos_importances, n_additional_best_snps, n_best_snps_per_donor):
    """snp_detection.py:218-227: the best positions per donor plus the best overall not among them."""
    importances_all = np.stack([imp for _chrom, _pos, imp, _base_counts in chrom_pos_importances], axis=0)
    best_snps_for_donors = np.argsort(-importances_all, axis=0)[:n_best_snps_per_donor]
    best_snps_overall = np.argsort(-importances_all.sum(axis=1))
    is_new_snps = ~np.isin(best_snps_overall, best_snps_for_donors)
    total_new_snps = np.cumsum(is_new_snps, axis=0)
    best_snps_overall = best_snps_overall[:np.searchsorted(total_new_snps, n_additional_best_snps, side='right')]
    selected_snp_ids = np.union1d(best_snps_for_donors.flatten(), best_snps_overall)
    return [chrom_pos_importances[i] for i in selected_snp_ids]


def _export_snps_to_beta(selected_snps, prior_filename):
    """snp_detection.py:230-242: an empty frame indexed by (CHROM, POS, BASE) -- positions known, betas unknown."""
    df = defaultdict(list)
    for chromosome, position, _importances, bases_count in selected_snps:
        for base, _base_count in bases_count.items():
            df['CHROM'].append(chromosome)
            df['POS'].append(position)
            df['BASE'].append(base)
    df = pd.DataFrame(df)
    df = df.set_index(['CHROM', 'POS', 'BASE'])
    df.to_parquet(prior_filename)


def detect_snps_positions(bamfile_location, genotypes: ProbabilisticGenotypes, barcode_handler: BarcodeHandler, *,
                          minimum_coverage: int, minimum_alternative_fraction: float = 0.01,
                          minimum_alternative_coverage: int = 100, n_best_snps_per_donor: int = 100,
                          n_additional_best_snps: int = 1000, regularization: float = 3.,
                          parse_read=cellranger_parse_read, joblib_n_jobs=-1, result_beta_prior_filename=None,
                          ignore_known_snps=True, max_fragment_step=10_000_000, joblib_verbosity=11):
    """
    Detects SNPs from the data, starting from loosely known genotypes (snp_detection.py:128-215); same arguments,
    same returned list of (chromosome, position, importance per donor, base counts).
    """
    from .demultiplexer import Demultiplexer
    # step 1: rough demultiplexing with the known genotypes
    snps = count_snps(bamfile_location=bamfile_location, chromosome2positions=genotypes.get_chromosome2positions(),
                      barcode_handler=barcode_handler, joblib_n_jobs=joblib_n_jobs, parse_read=parse_read,
                      joblib_verbosity=joblib_verbosity)
    _likelihoods, posterior_probabilities = Demultiplexer.predict_posteriors(
        snps, genotypes=genotypes, barcode_handler=barcode_handler, doublet_prior=0.0)
    barcode2donor = posterior_probabilities[posterior_probabilities.max(axis=1).gt(0.8)].idxmax(axis=1).to_dict()
    donor_counts = Counter(barcode2donor.values())
    print('Number of SNPs used for each donor during inference')
    print(pd.Series(donor_counts).sort_index())

    # step 2: collect SNPs using the predictions of the rough demultiplexing
    filename = bamfile_location if isinstance(bamfile_location, (str, Path)) else list(bamfile_location.values())[0]
    info = _bam_info(filename)  # header only: the alignments are streamed by the counting tasks, never held whole
    chromosomes = list(zip(info['names'], info['lengths']))
    sorted_donors = np.unique([donor for donor in barcode2donor.values()])
    tasks = [dict(bamfile_path=bamfile_location, chromosome=chromosome, start=start,
                  stop=min(start + max_fragment_step, length), barcode2donor=barcode2donor, parse_read=parse_read,
                  sorted_donors=sorted_donors, minimum_coverage=minimum_coverage,
                  minimum_alternative_coverage=minimum_alternative_coverage,
                  minimum_alternative_fraction=minimum_alternative_fraction, barcode_handler=barcode_handler,
                  regularization=regularization)
             for chromosome, length in chromosomes for start in range(0, length, max_fragment_step)]
    if joblib_n_jobs == 1 or len(tasks) <= 1:
        collection = [detect_snps_for_chromosome(**task) for task in tasks]
    else:
        import joblib
        with joblib.Parallel(n_jobs=joblib_n_jobs, verbose=joblib_verbosity, pre_dispatch='all') as parallel:
            collection = parallel(joblib.delayed(detect_snps_for_chromosome)(**task) for task in tasks)
    chrom_pos_importances = sum(collection, [])
    selected_snps = _select_top_snps(chrom_pos_importances, n_additional_best_snps, n_best_snps_per_donor)
    snp_positions = genotypes.get_snp_positions_set()
    if ignore_known_snps:
        selected_snps = [(chrom, pos, importance, base_count) for chrom, pos, importance, base_count in selected_snps
                         if (chrom, pos) not in snp_positions]
    if result_beta_prior_filename is not None:
        _export_snps_to_beta(selected_snps, result_beta_prior_filename)
    return selected_snps
