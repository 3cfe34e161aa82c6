"""
`detect_snps_positions` (demuxalot/snp_detection.py), SURVEY.md section 8(f) rank 2.

CPU: every stage of the wrapper against the unmodified reference (through the pysam stand-in) on the bundled example
BAM, when /root/reference is mounted; the coverage counter against a plain per-base loop on a synthetic BAM anywhere.
GPU: the whole call on a synthetic BAM, with the CUDA `predict_posteriors` inside, against the same glue driven by
the oracle's `predict_posteriors`.
"""
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from bam_writer import encode_read, write_bam
from demuxalot_b200 import BarcodeHandler, ProbabilisticGenotypes, snp_detection
from demuxalot_b200.counting import _open_cached, parse_read
from reference_loader import reference_available

EXAMPLE = Path('/root/reference/examples/example_data')


def make_pool_bam(path, seed=11, n_donors=4, n_barcodes=80, reads_per_barcode=60, length=1500):
    """A small pooled data set: donors differ at 60 known and 40 hidden SNPs of two contigs."""
    rng = np.random.default_rng(seed)
    contigs = [('chrA', length), ('chrB', length)]
    reference = {name: rng.integers(0, 4, size=n) for name, n in contigs}
    snps = {}  # (contig, pos) -> (ref base, alt base, dosage per donor), first 60 known
    while len(snps) < 100:
        name = contigs[int(rng.integers(0, 2))][0]
        pos = int(rng.integers(5, length - 5))
        ref_base = int(reference[name][pos])
        snps.setdefault((name, pos), (ref_base, (ref_base + int(rng.integers(1, 4))) % 4, rng.integers(0, 3, size=n_donors)))
    known = list(snps)[:60]
    donors = [f'Donor{k + 1}' for k in range(n_donors)]
    barcodes = sorted(''.join(rng.choice(list('ACGT'), size=12)) + '-1' for _ in range(n_barcodes))
    donor_of = rng.integers(0, n_donors, size=n_barcodes)
    records = []
    for b, barcode in enumerate(barcodes):
        for r in range(reads_per_barcode):
            ref_id = int(rng.integers(0, 2))
            name = contigs[ref_id][0]
            start = int(rng.integers(0, length - 50))
            seq = reference[name][start:start + 50].copy()
            for (snp_contig, pos), (_ref_base, alt_base, dosage) in snps.items():
                if snp_contig == name and start <= pos < start + 50 and rng.random() < dosage[donor_of[b]] / 2:
                    seq[pos - start] = alt_base
            qual = np.where(rng.random(50) < 0.05, 10, 32)
            tags = {'NH': 1, 'AS': 49, 'UB': ''.join(rng.choice(list('ACGT'), size=8)), 'CB': barcode}
            records.append((ref_id, start, encode_read(ref_id, start, [('M', 50)], ''.join('ACGT'[c] for c in seq),
                                                       qual.tolist(), mapq=255, name=f'r{b}_{r}', tags=tags)))
    records.sort(key=lambda t: (t[0], t[1]))
    write_bam(path, contigs, [rec for _ref, _pos, rec in records])
    genotypes = ProbabilisticGenotypes(donors)
    for (name, pos) in known:
        ref_base, alt_base, dosage = snps[(name, pos)]
        ref_row = genotypes.get_variant_id(name, pos, 'ACGT'[ref_base])
        alt_row = genotypes.get_variant_id(name, pos, 'ACGT'[alt_base])
        genotypes.variant_betas[ref_row] += 100 * (2 - dosage) / 2
        genotypes.variant_betas[alt_row] += 100 * dosage / 2
    return genotypes, BarcodeHandler(barcodes), snps, known


def test_count_coverage_against_a_plain_loop(tmp_path):
    import pysam_shim
    path = tmp_path / 'cov.bam'
    rng = np.random.default_rng(3)
    reads = []
    for k in range(300):  # every cigar operation, qualities around the threshold, N bases
        start = int(rng.integers(0, 900))
        cigar = [('S', 3), ('M', 12), ('I', 2), ('M', 7), ('D', 4), ('M', 9), ('N', 30), ('=', 5), ('X', 2), ('H', 4)]
        n_query = sum(l for op, l in cigar if op in 'MIS=X')
        seq = ''.join(rng.choice(list('ACGTN'), p=[.24, .24, .24, .24, .04], size=n_query))
        qual = rng.integers(5, 40, size=n_query).tolist()
        reads.append((start, encode_read(0, start, cigar, seq, qual, name=f'r{k}',
                                         tags={'NH': 1 + int(k % 7 == 0), 'AS': n_query - 1, 'UB': 'ACGT', 'CB': 'X-1'})))
    reads.sort(key=lambda t: t[0])
    write_bam(path, [('chr1', 1000)], [r for _s, r in reads])
    from demuxalot_b200.counting import count_coverage_native, native_io
    for start, stop in ((0, 1000), (137, 611)):
        mine = snp_detection.count_coverage(_open_cached(path), 'chr1', start, stop, lambda r: parse_read(r) is not None)
        with pysam_shim.AlignmentFile(str(path)) as f:
            want = np.asarray(f.count_coverage('chr1', start=start, stop=stop,
                                               read_callback=lambda r: parse_read(r) is not None), dtype='int32')
        assert mine.dtype == np.int32 and np.array_equal(mine, want) and mine.sum() > 1000
        if native_io() is not None:  # the native loop (built-in read filters) gives the same matrix
            fast = count_coverage_native(path, 'chr1', start, stop, parse_read)
            assert fast.dtype == np.int32 and np.array_equal(fast, want)
            assert count_coverage_native(path, 'chr1', start, stop, lambda r: parse_read(r)) is None  # custom callback


@pytest.mark.skipif(not reference_available(), reason='/root/reference not mounted')
def test_stages_match_the_reference_on_the_example_bam(tmp_path):
    import pysam_shim
    saved = sys.modules.get('pysam')
    sys.modules['pysam'] = pysam_shim
    try:
        from reference_loader import load_reference
        ref = load_reference()
        from demuxalot import snp_detection as ref_detection
        ref_detection.pysam = pysam_shim  # the module may have been imported with the bare stub before
        bam = str(EXAMPLE / 'test_bamfile.bam')
        handler = BarcodeHandler.from_file(EXAMPLE / 'test_barcodes.csv')
        ref_handler = ref.BarcodeHandler.from_file(EXAMPLE / 'test_barcodes.csv')
        barcode2donor = {bc: f'Donor0{1 + k % 4}' for k, bc in enumerate(handler.ordered_barcodes) if k % 5 != 0}
        kwargs = dict(chromosome='chr2', start=0, stop=1000, sorted_donors=np.unique(list(barcode2donor.values())),
                      barcode2donor=barcode2donor, regularization=3., minimum_coverage=30,
                      minimum_alternative_fraction=0.01, minimum_alternative_coverage=5)
        from demuxalot import snp_counter as ref_counter
        ref_counter.pysam = pysam_shim
        mine = snp_detection.detect_snps_for_chromosome(bam, parse_read=parse_read, barcode_handler=handler, **kwargs)
        want = ref_detection.detect_snps_for_chromosome(bam, parse_read=ref.cellranger_specific.parse_read,
                                                        barcode_handler=ref_handler, **kwargs)
    finally:
        if saved is not None:
            sys.modules['pysam'] = saved
    assert len(mine) == len(want) > 100
    for a, b in zip(mine, want):
        assert a[0] == b[0] and a[1] == b[1] and type(a[1]) is type(b[1])
        assert np.array_equal(a[2], b[2]) and list(a[3].items()) == list(b[3].items())
    top_mine = snp_detection._select_top_snps(mine, 20, 5)
    top_want = ref_detection._select_top_snps(want, 20, 5)
    assert [(c, p) for c, p, *_ in top_mine] == [(c, p) for c, p, *_ in top_want]
    snp_detection._export_snps_to_beta(top_mine, tmp_path / 'mine.parquet')
    ref_detection._export_snps_to_beta(top_want, tmp_path / 'want.parquet')
    assert (tmp_path / 'mine.parquet').read_bytes() == (tmp_path / 'want.parquet').read_bytes()
    # the exported positions import as new, genotype-free variants (the "detected SNVs" of BASELINE config 3)
    genotypes = ProbabilisticGenotypes(['Donor01', 'Donor02', 'Donor03', 'Donor04'])
    genotypes.add_prior_betas(tmp_path / 'mine.parquet', prior_strength=10)
    assert genotypes.n_variants == 2 * len(top_mine) and float(np.abs(genotypes.get_betas()).sum()) == 0.0


@pytest.mark.gpu
def test_detect_snps_positions_end_to_end_on_a_synthetic_pool(tmp_path, monkeypatch, native_lib):
    import oracle
    from demuxalot_b200 import Demultiplexer
    genotypes, handler, snps, known = make_pool_bam(tmp_path / 'pool.bam')
    kwargs = dict(minimum_coverage=20, minimum_alternative_coverage=4, n_best_snps_per_donor=6,
                  n_additional_best_snps=10, joblib_n_jobs=1, result_beta_prior_filename=tmp_path / 'new.parquet')
    got = snp_detection.detect_snps_positions(str(tmp_path / 'pool.bam'), genotypes, handler, **kwargs)
    monkeypatch.setattr(Demultiplexer, 'predict_posteriors',
                        staticmethod(lambda *a, **k: oracle.OracleDemultiplexer.predict_posteriors(*a, **k)))
    kwargs['result_beta_prior_filename'] = tmp_path / 'oracle.parquet'
    want = snp_detection.detect_snps_positions(str(tmp_path / 'pool.bam'), genotypes, handler, **kwargs)
    assert len(got) == len(want) > 5
    for a, b in zip(got, want):
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) and a[3] == b[3]
    assert (tmp_path / 'new.parquet').read_bytes() == (tmp_path / 'oracle.parquet').read_bytes()
    hidden = {key for key in snps if key not in set(known)}
    found = {(c, int(p)) for c, p, *_ in got}
    assert found and found <= hidden | set(known) and not (found & set(known))  # known positions are filtered out
    assert len(found & hidden) >= 5
