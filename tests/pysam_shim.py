"""
Test tooling: a small pure-Python stand-in for the parts of `pysam` that the reference's input stage touches
(`AlignmentFile.fetch / get_index_statistics / get_reference_length`, read attributes and tags, `VariantFile`).
pysam cannot be installed here (no wheel, no network); with this module registered as `sys.modules['pysam']` the
UNMODIFIED reference runs `add_vcf` + `count_snps` on its bundled example data, which is how
tests/golden/make_example_fixture.py produces the config #1 fixture.  Not used by the product.
"""
from __future__ import annotations

import gzip
import struct
from collections import namedtuple
from types import SimpleNamespace
from typing import Dict, Iterator, List

import numpy as np

_SEQ_CODE = '=ACMGRSVTWYHKDBN'
_SEQ_PAIRS = np.array([a + b for a in _SEQ_CODE for b in _SEQ_CODE])
_CONSUMES_REFERENCE = {0, 2, 3, 7, 8}


class AlignedSegment:
    __slots__ = ('reference_id', 'reference_start', 'reference_end', 'mapq', 'flag', 'cigartuples', 'seq',
                 'query_qualities', '_tags', 'query_name')

    @property
    def pos(self) -> int:
        return self.reference_start

    @property
    def mapping_quality(self) -> int:
        return self.mapq

    def has_tag(self, tag: str) -> bool:
        return tag in self._tags

    def get_tag(self, tag: str):
        return self._tags[tag]  # KeyError like pysam


AlignedRead = AlignedSegment
_TAG_FMT = {'c': '<b', 'C': '<B', 's': '<h', 'S': '<H', 'i': '<i', 'I': '<I', 'f': '<f'}


def _parse_tags(buf: bytes, off: int, end: int) -> Dict[str, object]:
    tags = {}
    while off < end:
        name = buf[off:off + 2].decode()
        kind = chr(buf[off + 2])
        off += 3
        if kind == 'A':
            tags[name] = chr(buf[off]); off += 1
        elif kind in _TAG_FMT:
            fmt = _TAG_FMT[kind]
            tags[name] = struct.unpack_from(fmt, buf, off)[0]; off += struct.calcsize(fmt)
        elif kind in 'ZH':
            stop = buf.index(b'\x00', off)
            tags[name] = buf[off:stop].decode(); off = stop + 1
        elif kind == 'B':
            sub = chr(buf[off]); count = struct.unpack_from('<i', buf, off + 1)[0]
            fmt = _TAG_FMT[sub]; size = struct.calcsize(fmt)
            tags[name] = [struct.unpack_from(fmt, buf, off + 5 + k * size)[0] for k in range(count)]
            off += 5 + count * size
        else:
            raise NotImplementedError(f'BAM tag type {kind!r}')
    return tags


class AlignmentFile:
    """Reads the whole (small) BAM into memory; `fetch` filters by overlap instead of using the .bai index."""

    def __init__(self, filename, mode='rb'):
        with gzip.open(str(filename), 'rb') as f:  # BGZF is a series of gzip members
            data = f.read()
        assert data[:4] == b'BAM\x01', 'not a BAM file'
        l_text = struct.unpack_from('<i', data, 4)[0]
        off = 8 + l_text
        n_ref = struct.unpack_from('<i', data, off)[0]
        off += 4
        self.references: List[str] = []
        self.lengths: List[int] = []
        for _ in range(n_ref):
            l_name = struct.unpack_from('<i', data, off)[0]
            self.references.append(data[off + 4:off + 4 + l_name - 1].decode())
            self.lengths.append(struct.unpack_from('<i', data, off + 4 + l_name)[0])
            off += 8 + l_name
        self._reads: Dict[int, List[AlignedSegment]] = {k: [] for k in range(n_ref)}
        self._mapped = [0] * n_ref
        while off < len(data):
            block_size = struct.unpack_from('<i', data, off)[0]
            rec_end = off + 4 + block_size
            (ref_id, pos, l_read_name, mapq, _bin, n_cigar, flag, l_seq, _nref, _npos, _tlen) = struct.unpack_from(
                '<iiBBHHHiiii', data, off + 4)
            p = off + 36
            read = AlignedSegment()
            read.query_name = data[p:p + l_read_name - 1].decode()
            p += l_read_name
            cigar = struct.unpack_from(f'<{n_cigar}I', data, p) if n_cigar else ()
            p += 4 * n_cigar
            read.cigartuples = [(c & 0xF, c >> 4) for c in cigar]
            packed = np.frombuffer(data, dtype=np.uint8, count=(l_seq + 1) // 2, offset=p)
            read.seq = ''.join(_SEQ_PAIRS[packed])[:l_seq]
            p += (l_seq + 1) // 2
            read.query_qualities = np.frombuffer(data, dtype=np.uint8, count=l_seq, offset=p)
            p += l_seq
            read._tags = _parse_tags(data, p, rec_end)
            read.reference_id, read.reference_start, read.mapq, read.flag = ref_id, pos, mapq, flag
            ref_len = sum(l for op, l in read.cigartuples if op in _CONSUMES_REFERENCE)
            read.reference_end = pos + ref_len if ref_len else None
            if ref_id >= 0:
                self._reads[ref_id].append(read)
                if not flag & 4:
                    self._mapped[ref_id] += 1
            off = rec_end
        for reads in self._reads.values():
            reads.sort(key=lambda r: r.reference_start)  # coordinate order (stable; already sorted files unchanged)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def get_index_statistics(self):
        Stat = namedtuple('IndexStats', ['contig', 'mapped', 'unmapped', 'total'])
        return [Stat(name, self._mapped[k], len(self._reads[k]) - self._mapped[k], len(self._reads[k]))
                for k, name in enumerate(self.references)]

    def get_reference_length(self, reference: str) -> int:
        return self.lengths[self.references.index(reference)]

    def count_coverage(self, contig, start=None, stop=None, quality_threshold=15, read_callback='all'):
        """Plain per-base loop with pysam's documented semantics (aligned bases, base quality >= threshold, ACGT)."""
        start = 0 if start is None else start
        stop = self.get_reference_length(contig) if stop is None else stop
        counts = {b: np.zeros(stop - start, dtype=np.int64) for b in 'ACGT'}
        for read in self.fetch(contig, start, stop):
            if read_callback == 'all':
                if read.flag & (0x4 | 0x100 | 0x200 | 0x400):
                    continue
            elif read_callback != 'nofilter' and not read_callback(read):
                continue
            qpos, rpos = 0, read.reference_start
            for op, length in read.cigartuples:
                if op in (0, 7, 8):
                    for k in range(length):
                        r = rpos + k
                        if start <= r < stop and read.query_qualities[qpos + k] >= quality_threshold:
                            base = read.seq[qpos + k]
                            if base in counts:
                                counts[base][r - start] += 1
                    qpos += length
                    rpos += length
                elif op in (1, 4):
                    qpos += length
                elif op in (2, 3):
                    rpos += length
        return tuple(counts[b] for b in 'ACGT')

    def fetch(self, contig=None, start=None, stop=None, **_kw) -> Iterator[AlignedSegment]:
        ids = range(len(self.references)) if contig is None else [self.references.index(contig)]
        for k in ids:
            for read in self._reads[k]:
                if read.flag & 4:
                    continue
                end = read.reference_end if read.reference_end is not None else read.reference_start + 1
                if start is not None and end <= start:
                    continue
                if stop is not None and read.reference_start >= stop:
                    continue
                yield read


class VariantFile:
    """Plain-text VCF records with the attributes genotypes.py:123-154 reads."""

    def __init__(self, filename, mode='r'):
        self._filename = str(filename)

    def fetch(self):
        opener = gzip.open if self._filename.endswith('.gz') else open
        samples: List[str] = []
        with opener(self._filename, 'rt') as f:
            for line in f:
                if line.startswith('##'):
                    continue
                fields = line.rstrip('\n').split('\t')
                if line.startswith('#CHROM'):
                    samples = fields[9:]
                    continue
                gt_slot = fields[8].split(':').index('GT')
                calls = {}
                for name, value in zip(samples, fields[9:]):
                    gt = value.split(':')[gt_slot].replace('|', '/').split('/')
                    calls[name] = {'GT': tuple(None if a == '.' else int(a) for a in gt)}
                alts = tuple(a for a in fields[4].split(',') if a != '.')
                yield SimpleNamespace(chrom=fields[0], pos=int(fields[1]), alleles=(fields[3],) + alts, samples=calls)
