"""
Input stage without pysam: `demuxalot_b200.bam` + `demuxalot_b200.counting.count_snps`.

* offline: a synthetic BAM written by tests/bam_writer.py exercises the BGZF/BAM parser, region fetch, cigar
  handling, read filters, UMI grouping and the per-position base consensus against hand-computed expectations;
* with /root/reference mounted: `count_snps` on the reference's bundled example BAM must reproduce, element for
  element, the calls the UNMODIFIED reference produced (stored subset fixture; the full live comparison runs with
  DMX_SLOW_TESTS=1).
"""
import os
from pathlib import Path

import numpy as np
import pytest

from bam_writer import encode_read, write_bam
from demuxalot_b200 import BarcodeHandler, ProbabilisticGenotypes
from demuxalot_b200.bam import BamFile
from demuxalot_b200.counting import SnpPositions, collapse_molecule, count_snps, hash_string, parse_read
from golden_io import GOLDEN_DIR
from reference_loader import reference_available

EXAMPLE = Path('/root/reference/examples/example_data')


def _read(pos, cigar, seq, cb='AAA-1', ub='ACGT', nh=1, ascore=None, mapq=255, q=30, ref_id=0, name='r'):
    ascore = len(seq) - 2 if ascore is None else ascore
    tags = {'NH': nh, 'AS': ascore}
    if ub is not None:
        tags['UB'] = ub
    if cb is not None:
        tags['CB'] = cb
    return encode_read(ref_id, pos, cigar, seq, [q] * len(seq), mapq=mapq, name=name, tags=tags)


def test_hash_string_matches_reference_formula():
    assert hash_string('A') == 65 and hash_string('AC') == 65 * 5 + 67
    assert hash_string('ACGTACGTACGT') == sum(ord(c) * 5 ** (11 - k) for k, c in enumerate('ACGTACGTACGT')) % 2147483629


def test_bam_reader_roundtrip_and_fetch(tmp_path):
    reads = [
        _read(100, [('M', 10)], 'ACGTACGTAC', name='a'),
        _read(105, [('S', 2), ('M', 4), ('N', 100), ('M', 4)], 'TTACGTACGT', name='b', q=12),
        _read(300, [('M', 3), ('I', 2), ('M', 3), ('D', 4), ('M', 2)], 'ACGTTACGTA', name='c'),
        encode_read(0, 400, [], 'ACGT', [1, 2, 3, 4], flag=4, name='unmapped'),
        _read(50, [('M', 5)], 'GGGGG', ref_id=1, name='d'),
    ]
    path = tmp_path / 't.bam'
    write_bam(path, [('chr1', 1000), ('chr2', 500)], reads, block_bytes=97)
    bam = BamFile(path)
    assert bam.references == ['chr1', 'chr2'] and bam.lengths == [1000, 500]
    assert bam.mapped_reads_per_reference() == {'chr1': 3, 'chr2': 1}
    assert bam.get_reference_length('chr2') == 500
    got = list(bam.fetch('chr1'))
    assert [r.query_name for r in got] == ['a', 'b', 'c']
    a, b, c = got
    assert (a.reference_start, a.reference_end, a.seq, list(a.query_qualities)) == (100, 110, 'ACGTACGTAC', [30] * 10)
    assert b.cigartuples == [(4, 2), (0, 4), (3, 100), (0, 4)] and b.reference_end == 105 + 108
    assert c.reference_end == 300 + 3 + 3 + 4 + 2 and c.mapq == 255 and c.pos == 300
    assert a.get_tag('UB') == 'ACGT' and a.get_tag('NH') == 1 and a.has_tag('CB') and not a.has_tag('XX')
    with pytest.raises(KeyError):
        a.get_tag('XX')
    assert [r.query_name for r in bam.fetch('chr1', start=110, stop=300)] == ['b']       # half-open overlap
    assert [r.query_name for r in bam.fetch('chr1', start=109, stop=301)] == ['a', 'b', 'c']
    assert [r.query_name for r in bam.fetch('chr2', 0, 51)] == ['d'] and list(bam.fetch('chr2', 0, 50)) == []


def test_snp_positions_and_cigar_walk(tmp_path):
    snps = SnpPositions(np.array([102, 107, 210, 306, 313]))
    assert snps.any_in(100, 103) and not snps.any_in(103, 107) and snps.any_in(107, 108) and not snps.any_in(314, 999)
    path = tmp_path / 't.bam'
    write_bam(path, [('chr1', 1000)], [
        _read(105, [('S', 2), ('M', 4), ('N', 100), ('M', 4)], 'TTACGTACGT'),
        _read(300, [('M', 3), ('I', 2), ('M', 3), ('D', 4), ('M', 2)], 'ACGTTACGTA'),
    ])
    b, c = list(BamFile(path).fetch('chr1'))
    # b: soft clip 2, M4 covers 105..108 (read 2..5), skip 100, M4 covers 209..212 (read 6..9)
    assert snps.calls_of_read(b) == [(107, 'G', 30), (210, 'C', 30)]
    # c: M3 300..302, I2, M3 303..305 (read 5..7), D4 306..309, M2 310..311 -> SNPs 306 (deleted) and 313: none
    assert snps.calls_of_read(c) == []


def test_collapse_molecule_rules():
    class R:
        def __init__(self, start, end, ascore, calls):
            self.reference_start, self.reference_end, self._as, self.calls = start, end, ascore, calls
        def get_tag(self, _t): return self._as

    class Lookup:
        def calls_of_read(self, read): return read.calls

    r1 = R(0, 50, 48, [(10, 'A', 30), (20, 'C', 50), (30, 'G', 10)])
    dup = R(0, 50, 48, [(10, 'T', 30)])                      # same (start, end, AS): ignored entirely
    r2 = R(5, 55, 50, [(10, 'A', 20), (20, 'T', 9), (30, 'T', 10)])
    p_group, calls = collapse_molecule([(r1, 0.01), (dup, 0.5), (r2, 0.02)], Lookup())
    assert p_group == 0.01 * 0.02
    got = {pos: (base, p) for pos, base, p in calls}
    assert got[10] == ('A', 1 * 0.1 ** (0.1 * 30) * 0.1 ** (0.1 * 20))          # same base: product
    assert got[20] == ('C', 0.1 ** (0.1 * 40))                                   # q capped at 40; T (q9) is >1000x worse
    assert 30 not in got                                                         # two equally bad candidates: no call


@pytest.fixture(scope='module', autouse=True)
def _host_library():
    from demuxalot_b200 import build
    build.build_host()


def test_native_library_exports_every_declared_symbol():
    import re
    from demuxalot_b200.counting import native_io
    header = (Path(__file__).resolve().parent.parent / 'include' / 'demux_io.h').read_text()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(dmxio_[a-z0-9_]+)\s*\(', header))
    lib = native_io()
    assert lib is not None and declared
    for name in declared:
        assert hasattr(lib, name), name


@pytest.mark.parametrize('use_native', [True, False], ids=['native', 'python'])
def test_count_snps_on_synthetic_bam(tmp_path, use_native):
    genotypes = ProbabilisticGenotypes(['D1', 'D2'])
    for pos, bases in ((102, 'AC'), (107, 'CT'), (2100, 'GT'), (3004, 'AC')):
        for b in bases:
            genotypes.get_variant_id('chr1', pos, b)
    handler = BarcodeHandler(['AAA-1', 'CCC-1'])
    reads = [
        _read(100, [('M', 10)], 'ACATACGCAC', cb='CCC-1', ub='AAAA'),            # 102 -> A, 107 -> C
        _read(100, [('M', 10)], 'ACATACGCAC', cb='CCC-1', ub='AAAA'),            # complete duplicate: ignored
        _read(101, [('M', 10)], 'CATACGTACG', cb='CCC-1', ub='AAAA', q=20),      # same molecule: 102 -> A, 107 -> T
        _read(100, [('M', 10)], 'ACCTACGTAC', cb='AAA-1', ub='CCCC'),            # 102 -> C, 107 -> T
        _read(100, [('M', 10)], 'ACCTACGTAC', cb='GGG-1', ub='CCCC'),            # barcode not whitelisted
        _read(100, [('M', 10)], 'ACCTACGTAC', cb='AAA-1', ub=None),              # no UMI
        _read(100, [('M', 10)], 'ACCTACGTAC', cb='AAA-1', ub='TTTT', nh=2),      # multi-mapped
        _read(100, [('M', 10)], 'ACCTACGTAC', cb='AAA-1', ub='TTTT', ascore=1),  # too many edits
        _read(100, [('M', 10)], 'ACCTACGTAC', cb='AAA-1', ub='TTTT', mapq=3),    # low mapq
        # joins the still-open AAA-1/CCCC group (groups are only closed 1000 bp after their furthest read end,
        # and the check runs after the read was added), then closes CCC-1/AAAA: 2100 -> G
        _read(2095, [('M', 10)], 'ACGTAGACGT', cb='AAA-1', ub='CCCC'),
        _read(3000, [('M', 10)], 'ACGTCGACGT', cb='CCC-1', ub='AAAA'),           # CCC-1/AAAA again: a new molecule
    ]
    path = tmp_path / 't.bam'
    write_bam(path, [('chr1', 5000)], reads)
    calls = count_snps(str(path), genotypes.get_chromosome2positions(), handler, joblib_n_jobs=1,
                       use_native=use_native)
    assert list(calls) == ['chr1']
    c = calls['chr1']
    assert c.n_molecules == 3 and c.n_snp_calls == 5
    mol = c.molecules[:3]
    assert list(mol['compressed_cb']) == [1, 0, 1]
    assert list(mol['compressed_ub']) == [hash_string('AAAA'), hash_string('CCCC'), hash_string('AAAA')]
    assert np.array_equal(mol['p_group_misaligned'], np.float32([0.01 * 0.01, 0.01 * 0.01, 0.01]))
    sc = c.snp_calls[:5]
    assert list(sc['molecule_index']) == [0, 1, 1, 1, 2]
    assert list(sc['snp_position']) == [102, 102, 107, 2100, 3004]
    assert list(sc['base_index']) == [0, 1, 3, 2, 1]                             # A | C, T, G | C
    q30, q20 = 0.1 ** (0.1 * 30), 0.1 ** (0.1 * 20)
    # molecule 0, position 102: A seen twice -> product; position 107: C (q30) vs T (q20) differ by only 10x, both
    # survive the 1000x rule, so the position is ambiguous and yields no call
    assert np.array_equal(sc['p_base_wrong'], np.float32([1 * q30 * q20, q30, q30, q30, q30]))


def test_native_and_python_loops_agree_on_a_random_bam(tmp_path):
    """Randomised reads (soft clips, insertions, deletions, skips, duplicates, filtered reads, several barcodes and
    UMIs, two references) through both implementations: identical records."""
    rng = np.random.default_rng(11)
    barcodes = [f'BC{k:03d}-1' for k in range(12)]
    umis = [''.join(rng.choice(list('ACGT'), size=8)) for _ in range(40)]
    reads = []
    for ref_id in (0, 1):
        pos = 0
        for _ in range(1500):
            pos += int(rng.integers(0, 9))
            ops, length = [], 0
            if rng.random() < 0.2:
                ops.append(('S', int(rng.integers(1, 5))))
            ops.append(('M', int(rng.integers(8, 30))))
            extra = rng.random()
            if extra < 0.15:
                ops += [('I', int(rng.integers(1, 4))), ('M', int(rng.integers(5, 20)))]
            elif extra < 0.3:
                ops += [('D', int(rng.integers(1, 6))), ('M', int(rng.integers(5, 20)))]
            elif extra < 0.45:
                ops += [('N', int(rng.integers(50, 1500))), ('M', int(rng.integers(5, 20)))]
            if rng.random() < 0.1:
                ops.append(('H', 3))
            n_query = sum(l for op, l in ops if op in 'MIS=X')
            seq = ''.join(rng.choice(list('ACGTN'), size=n_query, p=[.24, .24, .24, .24, .04]))
            qual = rng.integers(2, 42, size=n_query).tolist()
            tags = {'NH': int(rng.choice([1, 1, 1, 2])), 'AS': int(n_query - rng.integers(0, 12)),
                    'CB': str(rng.choice(barcodes + ['ZZZ-1']))}
            if rng.random() < 0.95:
                tags['UB'] = str(rng.choice(umis))
            reads.append(encode_read(ref_id, pos, ops, seq, qual, mapq=int(rng.choice([255, 255, 3])), tags=tags))
            if rng.random() < 0.1:
                reads.append(reads[-1])  # complete duplicate
    path = tmp_path / 'r.bam'
    write_bam(path, [('chrA', 40000), ('chrB', 40000)], reads, block_bytes=1500)
    positions = {c: np.unique(rng.integers(0, 9000, size=700)) for c in ('chrA', 'chrB')}
    handler = BarcodeHandler(barcodes)
    native = count_snps(str(path), positions, handler, joblib_n_jobs=1, use_native=True)
    python = count_snps(str(path), positions, handler, joblib_n_jobs=1, use_native=False)
    assert list(native) == list(python)
    # region tasks on a thread pool (the native loop runs outside the GIL): same records, same order
    threaded = count_snps(str(path), positions, handler, joblib_n_jobs=2, use_native=True)
    assert list(threaded) == list(native)
    for chrom in native:
        assert np.array_equal(threaded[chrom].molecules, native[chrom].molecules)
        assert np.array_equal(threaded[chrom].snp_calls, native[chrom].snp_calls)
    total = 0
    for chrom in python:
        a, b = native[chrom], python[chrom]
        assert (a.n_molecules, a.n_snp_calls) == (b.n_molecules, b.n_snp_calls)
        assert np.array_equal(a.molecules[:a.n_molecules], b.molecules[:b.n_molecules])
        assert np.array_equal(a.snp_calls[:a.n_snp_calls], b.snp_calls[:b.n_snp_calls])
        total += a.n_snp_calls
    assert total > 500


@pytest.mark.skipif(not reference_available(), reason='/root/reference not mounted')
@pytest.mark.parametrize('use_native', [True, False], ids=['native', 'python'])
def test_example_bam_matches_reference_fixture(use_native):
    """Our reader + counting on the bundled example reproduce the reference's calls exactly (first 48 barcodes are
    stored in the committed fixture; sizes of the full run in example_data_summary.json)."""
    import json
    from bench import slice_barcodes
    from types import SimpleNamespace
    genotypes = ProbabilisticGenotypes(['Donor01', 'Donor02', 'Donor03', 'Donor04'])
    genotypes.add_vcf(EXAMPLE / 'test_genotypes.vcf')
    handler = BarcodeHandler.from_file(EXAMPLE / 'test_barcodes.csv')
    calls = count_snps(str(EXAMPLE / 'test_bamfile.bam'), genotypes.get_chromosome2positions(), handler, joblib_n_jobs=1,
                       use_native=use_native)
    summary = json.loads((GOLDEN_DIR / 'example_data_summary.json').read_text())
    assert list(calls) == list(summary['chromosomes'])  # same task order -> same dict order
    for chrom, sizes in summary['chromosomes'].items():
        assert calls[chrom].n_molecules == sizes['n_molecules'] and calls[chrom].n_snp_calls == sizes['n_snp_calls']
    sub, _ = slice_barcodes(SimpleNamespace(calls=calls, barcode_handler=handler), 48)
    fx = np.load(GOLDEN_DIR / 'example_data_48bc.npz')
    for chrom in calls:
        n_mol, n_calls = (int(x) for x in fx[f'n__{chrom}'])
        assert np.array_equal(sub[chrom].molecules[:sub[chrom].n_molecules], fx[f'mol__{chrom}'][:n_mol])
        assert np.array_equal(sub[chrom].snp_calls[:sub[chrom].n_snp_calls], fx[f'calls__{chrom}'][:n_calls])


@pytest.mark.skipif(not (reference_available() and os.environ.get('DMX_SLOW_TESTS')), reason='slow live comparison')
def test_example_bam_matches_live_reference():
    import subprocess, sys
    code = Path(__file__).with_name('live_count_snps_check.py')
    assert subprocess.run([sys.executable, str(code)], capture_output=True, text=True).stdout.strip().endswith('IDENTICAL')
