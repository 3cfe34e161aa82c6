"""Test tooling: writes small coordinate-sorted BAM files (BGZF + BAM records) so the reader can be tested offline."""
import struct
import zlib

_SEQ_CODE = '=ACMGRSVTWYHKDBN'
_CIGAR_OPS = 'MIDNSHP=X'


def _bgzf_block(payload: bytes) -> bytes:
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    data = comp.compress(payload) + comp.flush()
    size = 18 + len(data) + 8
    header = b'\x1f\x8b\x08\x04' + b'\x00' * 4 + b'\x00\xff' + struct.pack('<H', 6) + b'BC' + struct.pack('<HH', 2, size - 1)
    return header + data + struct.pack('<II', zlib.crc32(payload) & 0xFFFFFFFF, len(payload))


def _tag(name, value):
    if isinstance(value, str):
        return name.encode() + b'Z' + value.encode() + b'\x00'
    if isinstance(value, float):
        return name.encode() + b'f' + struct.pack('<f', value)
    if 0 <= value < 256:
        return name.encode() + b'C' + struct.pack('<B', value)
    return name.encode() + b'i' + struct.pack('<i', value)


def encode_read(ref_id, pos, cigar, seq, qual, mapq=255, flag=0, name='r', tags=None):
    """cigar: list of (op_char, length), e.g. [('M', 50), ('N', 200), ('M', 48)]."""
    name_b = name.encode() + b'\x00'
    cig = b''.join(struct.pack('<I', (l << 4) | _CIGAR_OPS.index(op)) for op, l in cigar)
    codes = [_SEQ_CODE.index(c) for c in seq] + [0]
    packed = bytes((codes[2 * k] << 4) | codes[2 * k + 1] for k in range((len(seq) + 1) // 2))
    tag_b = b''.join(_tag(k, v) for k, v in (tags or {}).items())
    core = struct.pack('<iiBBHHHiiii', ref_id, pos, len(name_b), mapq, 0, len(cigar), flag, len(seq), -1, -1, 0)
    body = core + name_b + cig + packed + bytes(qual) + tag_b
    return struct.pack('<i', len(body)) + body


def write_bam(path, references, reads, block_bytes=4096):
    """references: [(name, length)]; reads: encoded records, already in coordinate order."""
    text = ''.join(f'@SQ\tSN:{n}\tLN:{l}\n' for n, l in references).encode()
    data = b'BAM\x01' + struct.pack('<i', len(text)) + text + struct.pack('<i', len(references))
    for n, l in references:
        data += struct.pack('<i', len(n) + 1) + n.encode() + b'\x00' + struct.pack('<i', l)
    data += b''.join(reads)
    with open(path, 'wb') as f:
        for off in range(0, len(data), block_bytes):  # records deliberately straddle block boundaries
            f.write(_bgzf_block(data[off:off + block_bytes]))
        f.write(_bgzf_block(b''))  # EOF marker
