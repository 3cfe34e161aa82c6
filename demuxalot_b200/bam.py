"""
Minimal pysam-free BAM access for the input stage (`count_snps`): BGZF inflate with zlib, record parsing with
`struct`, lazy decoding of sequence / qualities / tags.  Only what the counting stage touches is implemented
(SURVEY.md section 8(f), rank 1): coordinate-ordered iteration over the reads of one reference that overlap a
region, per-reference mapped-read counts and reference lengths.

The whole file is inflated into memory once per process (fine for the bundled example and for lane-sized BAMs on a
GPU host; a streaming reader over the .bai index is the natural next step).
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path
from typing import Dict, Iterator, List, Optional

import numpy as np

_SEQ_CODE = '=ACMGRSVTWYHKDBN'
_PAIR_TABLE = [a + b for a in _SEQ_CODE for b in _SEQ_CODE]
_CONSUMES_REFERENCE = (0, 2, 3, 7, 8)
_TAG_FMT = {'c': ('<b', 1), 'C': ('<B', 1), 's': ('<h', 2), 'S': ('<H', 2), 'i': ('<i', 4), 'I': ('<I', 4), 'f': ('<f', 4)}
_CORE = struct.Struct('<iiBBHHHiiii')


def inflate_bgzf(path) -> bytes:
    """All BGZF members of a file, concatenated."""
    raw = Path(path).read_bytes()
    view = memoryview(raw)
    out = []
    off = 0
    while off + 18 <= len(raw):
        if raw[off] != 0x1F or raw[off + 1] != 0x8B:
            raise ValueError('not a BGZF stream')
        xlen = struct.unpack_from('<H', raw, off + 10)[0]
        block_size, p, extra_end = None, off + 12, off + 12 + xlen
        while p + 4 <= extra_end:  # extra subfields: the 'BC' one holds the total block size - 1
            si1, si2, slen = raw[p], raw[p + 1], struct.unpack_from('<H', raw, p + 2)[0]
            if si1 == 66 and si2 == 67:
                block_size = struct.unpack_from('<H', raw, p + 4)[0] + 1
            p += 4 + slen
        if block_size is None:
            raise ValueError('gzip member without BGZF block size')
        out.append(zlib.decompress(view[extra_end:off + block_size - 8], -15))  # raw deflate payload
        off += block_size
    return b''.join(out)


class BamRecord:
    """One alignment; attribute names follow pysam so that user `parse_read` callbacks keep working."""
    __slots__ = ('_buf', '_off', '_end', 'reference_id', 'reference_start', 'mapq', 'flag', '_l_read_name', '_n_cigar',
                 '_l_seq', '_cigar', '_ref_end', '_seq', '_qual', '_tags')

    def __init__(self, buf: bytes, off: int):
        block_size = struct.unpack_from('<i', buf, off)[0]
        (self.reference_id, self.reference_start, self._l_read_name, self.mapq, _bin, self._n_cigar, self.flag,
         self._l_seq, _nref, _npos, _tlen) = _CORE.unpack_from(buf, off + 4)
        self._buf, self._off, self._end = buf, off + 36, off + 4 + block_size
        self._cigar = self._ref_end = self._seq = self._qual = self._tags = None

    # -- pysam-compatible surface -------------------------------------------------------------------------
    @property
    def pos(self) -> int:
        return self.reference_start

    @property
    def mapping_quality(self) -> int:
        return self.mapq

    @property
    def query_name(self) -> str:
        return self._buf[self._off:self._off + self._l_read_name - 1].decode()

    @property
    def cigartuples(self):
        if self._cigar is None:
            p = self._off + self._l_read_name
            raw = struct.unpack_from(f'<{self._n_cigar}I', self._buf, p) if self._n_cigar else ()
            self._cigar = [(c & 0xF, c >> 4) for c in raw]
        return self._cigar

    @property
    def reference_end(self) -> Optional[int]:
        if self._ref_end is None:
            span = 0
            for op, length in self.cigartuples:
                if op in _CONSUMES_REFERENCE:
                    span += length
            # htslib's bam_endpos: a CIGAR that consumes no reference (all soft clip / insertion) ends at pos + 1;
            # None only for reads without a CIGAR (pysam also returns None for unmapped reads)
            if span:
                self._ref_end = self.reference_start + span
            else:
                self._ref_end = self.reference_start + 1 if len(self.cigartuples) else -1
        return None if self._ref_end < 0 else self._ref_end

    @property
    def seq(self) -> str:
        if self._seq is None:
            p = self._off + self._l_read_name + 4 * self._n_cigar
            packed = self._buf[p:p + (self._l_seq + 1) // 2]
            self._seq = ''.join([_PAIR_TABLE[b] for b in packed])[:self._l_seq]
        return self._seq

    query_sequence = seq

    @property
    def query_qualities(self):
        if self._qual is None:
            p = self._off + self._l_read_name + 4 * self._n_cigar + (self._l_seq + 1) // 2
            self._qual = self._buf[p:p + self._l_seq]  # bytes: indexing yields ints
        return self._qual

    def _all_tags(self) -> Dict[str, object]:
        if self._tags is None:
            buf = self._buf
            off = self._off + self._l_read_name + 4 * self._n_cigar + (self._l_seq + 1) // 2 + self._l_seq
            tags = {}
            while off < self._end:
                name = buf[off:off + 2].decode()
                kind = chr(buf[off + 2])
                off += 3
                if kind == 'Z' or kind == 'H':
                    stop = buf.index(b'\x00', off)
                    tags[name] = buf[off:stop].decode()
                    off = stop + 1
                elif kind == 'A':
                    tags[name] = chr(buf[off])
                    off += 1
                elif kind in _TAG_FMT:
                    fmt, size = _TAG_FMT[kind]
                    tags[name] = struct.unpack_from(fmt, buf, off)[0]
                    off += size
                elif kind == 'B':
                    fmt, size = _TAG_FMT[chr(buf[off])]
                    count = struct.unpack_from('<i', buf, off + 1)[0]
                    tags[name] = list(struct.unpack_from(f'<{count}{fmt[1]}', buf, off + 5))
                    off += 5 + count * size
                else:
                    raise ValueError(f'unknown BAM tag type {kind!r}')
            self._tags = tags
        return self._tags

    def has_tag(self, tag: str) -> bool:
        return tag in self._all_tags()

    def get_tag(self, tag: str):
        return self._all_tags()[tag]  # KeyError when absent, as pysam raises


class BamFile:
    def __init__(self, filename):
        data = inflate_bgzf(filename)
        if data[:4] != b'BAM\x01':
            raise ValueError(f'{filename} is not a BAM file')
        self._data = data
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
        # one pass over the record headers: offset, reference, position, flag of every alignment
        offsets, ref_ids, starts, flags = [], [], [], []
        n = len(data)
        while off < n:
            block_size, ref_id, pos = struct.unpack_from('<iii', data, off)
            offsets.append(off)
            ref_ids.append(ref_id)
            starts.append(pos)
            flags.append(struct.unpack_from('<H', data, off + 18)[0])
            off += 4 + block_size
        self._offsets = np.asarray(offsets, dtype=np.int64)
        self._ref_ids = np.asarray(ref_ids, dtype=np.int32)
        self._starts = np.asarray(starts, dtype=np.int64)
        self._flags = np.asarray(flags, dtype=np.uint16)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def get_reference_length(self, reference: str) -> int:
        return self.lengths[self.references.index(reference)]

    def mapped_reads_per_reference(self) -> Dict[str, int]:
        """What pysam's get_index_statistics() reports as `.mapped` per contig."""
        mapped = (self._flags & 4) == 0
        return {name: int(np.count_nonzero(mapped & (self._ref_ids == k))) for k, name in enumerate(self.references)}

    def fetch(self, contig: str, start: Optional[int] = None, stop: Optional[int] = None) -> Iterator[BamRecord]:
        """Mapped reads of `contig` overlapping [start, stop), in file (coordinate) order."""
        ref_id = self.references.index(contig)
        selected = np.flatnonzero((self._ref_ids == ref_id) & ((self._flags & 4) == 0))
        if stop is not None:
            selected = selected[self._starts[selected] < stop]
        data = self._data
        for off in self._offsets[selected].tolist():
            read = BamRecord(data, off)
            if start is not None:
                end = read.reference_end
                if (end if end is not None else read.reference_start + 1) <= start:
                    continue
            yield read


# ---------------------------------------------------------------------------------------------------------------
# header / index access without inflating the whole file (used by the native counting path)
# ---------------------------------------------------------------------------------------------------------------

def read_bam_header(path):
    """(reference names, reference lengths, BGZF virtual offset of the first alignment record)."""
    blocks: List[bytes] = []       # inflated blocks read so far
    starts: List[int] = []         # compressed file offset of each block
    with open(path, 'rb') as f:
        def more() -> bool:
            coffset = f.tell()
            head = f.read(12)
            if len(head) < 12:
                return False
            xlen = struct.unpack_from('<H', head, 10)[0]
            extra = f.read(xlen)
            size, p = None, 0
            while p + 4 <= xlen:
                slen = struct.unpack_from('<H', extra, p + 2)[0]
                if extra[p] == 66 and extra[p + 1] == 67:
                    size = struct.unpack_from('<H', extra, p + 4)[0] + 1
                p += 4 + slen
            if size is None:
                raise ValueError('not a BGZF stream')
            payload = f.read(size - 12 - xlen)
            blocks.append(zlib.decompress(payload[:-8], -15))
            starts.append(coffset)
            return True

        def need(n: int) -> bytes:
            while sum(len(b) for b in blocks) < n:
                if not more():
                    raise ValueError('truncated BAM header')
            return b''.join(blocks)

        data = need(12)
        if data[:4] != b'BAM\x01':
            raise ValueError(f'{path} is not a BAM file')
        l_text = struct.unpack_from('<i', data, 4)[0]
        data = need(12 + l_text)
        n_ref = struct.unpack_from('<i', data, 8 + l_text)[0]
        off = 12 + l_text
        names, lengths = [], []
        for _ in range(n_ref):
            data = need(off + 4)
            l_name = struct.unpack_from('<i', data, off)[0]
            data = need(off + 8 + l_name)
            names.append(data[off + 4:off + 4 + l_name - 1].decode())
            lengths.append(struct.unpack_from('<i', data, off + 4 + l_name)[0])
            off += 8 + l_name
        # translate the uncompressed offset of the first record into a virtual offset
        consumed = 0
        for block, coffset in zip(blocks, starts):
            if off < consumed + len(block) or (off == consumed + len(block) and block is blocks[-1]):
                if off == consumed + len(block):  # record starts exactly at the next block
                    more_ok = more()
                    return names, lengths, ((starts[-1] << 16) if more_ok else (coffset << 16) | len(block))
                return names, lengths, (coffset << 16) | (off - consumed)
            consumed += len(block)
        raise ValueError('could not locate the first alignment record')


def read_bai(path):
    """Per reference: dict(mapped, unmapped, begin_voffset, linear) from a .bai index (SAM spec section 5.2)."""
    raw = Path(path).read_bytes()
    if raw[:4] != b'BAI\x01':
        raise ValueError(f'{path} is not a BAI index')
    n_ref = struct.unpack_from('<i', raw, 4)[0]
    off = 8
    out = []
    for _ in range(n_ref):
        n_bin = struct.unpack_from('<i', raw, off)[0]
        off += 4
        entry = dict(mapped=0, unmapped=0, begin_voffset=None, linear=None)
        for _b in range(n_bin):
            bin_id, n_chunk = struct.unpack_from('<Ii', raw, off)
            off += 8
            if bin_id == 37450 and n_chunk == 2:  # metadata pseudo-bin
                beg, _end, mapped, unmapped = struct.unpack_from('<QQQQ', raw, off)
                entry.update(mapped=int(mapped), unmapped=int(unmapped), begin_voffset=int(beg))
            off += 16 * n_chunk
        n_intv = struct.unpack_from('<i', raw, off)[0]
        off += 4
        entry['linear'] = np.frombuffer(raw, dtype='<u8', count=n_intv, offset=off).copy()
        off += 8 * n_intv
        out.append(entry)
    return out


def region_start_voffset(index_entry: dict, start: Optional[int]) -> Optional[int]:
    """A virtual offset at or before the first read overlapping `start` (16 kb linear index), None if unknown."""
    linear = index_entry['linear']
    if start is not None and len(linear):
        window = min(int(start) >> 14, len(linear) - 1)
        nonzero = linear[:window + 1][linear[:window + 1] != 0]
        if len(nonzero) and linear[window] != 0:
            return int(linear[window])
        later = linear[window:][linear[window:] != 0]
        if len(later):
            return int(min(later[0], nonzero[-1]) if len(nonzero) else later[0])
    return index_entry['begin_voffset']
