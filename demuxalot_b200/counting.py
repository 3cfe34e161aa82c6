"""
Input stage: `count_snps` without pysam.

Same contract as the reference's `count_snps` (demuxalot/snp_counter.py:279-327): reads of one BAM are filtered by a
`parse_read` callback, grouped by (cell barcode, UMI), every group is collapsed to at most one base call per SNP
position (snp_counter.py:142-192) and appended to a per-chromosome `CompressedSNPCalls`.  Region tasks, their order
and the grouping rules follow the reference so that the output arrays are identical element for element
(tests/test_counting.py checks this against the unmodified reference running on a pysam stand-in).
The BAM itself is read by `demuxalot_b200.bam` (zlib + struct), so the stage runs where pysam is unavailable.
"""
from __future__ import annotations

import os

from collections import defaultdict
from pathlib import Path
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .bam import BamFile, BamRecord, read_bai, read_bam_header, region_start_voffset
from .barcodes import BarcodeHandler
from .calls import MOLECULE_DTYPE, SNP_CALL_DTYPE, CompressedSNPCalls

SEGMENT_LENGTH = 1000  # snp_counter.py:231: groups are flushed once reads have moved this far past them


def hash_string(s: str) -> int:
    """UMI string -> int, base-5 polynomial modulo the prime 2147483629 (demuxalot/utils.py:12-22)."""
    value = 0
    for ch in s:
        value = value * 5 + ord(ch)
    return value % 2147483629


def parse_read(read, umi_tag='UB', nhits_tag='NH', score_tag='AS', score_diff_max=8, mapq_threshold=20,
               p_misaligned_default=0.01) -> Optional[Tuple[float, int]]:
    """cellranger-flavoured read filter (demuxalot/cellranger_specific.py:13-36): None = skip the read."""
    if read.get_tag(score_tag) <= len(read.seq) - score_diff_max:
        return None  # too many edits
    if read.get_tag(nhits_tag) > 1:
        return None  # multi-mapped
    if not read.has_tag(umi_tag):
        return None
    if read.mapq < mapq_threshold:
        return None
    return p_misaligned_default, hash_string(read.get_tag(umi_tag))


def parse_read_bd_rhapsody(read, umi_tag='MA', score_tag='AS', score_diff_max=8, mapq_threshold=20,
                           p_misaligned_default=0.01) -> Optional[Tuple[float, int]]:
    """BD Rhapsody flavour (demuxalot/BDRhapsody_specific.py:13-36): no NH tag, UMI in MA."""
    if read.get_tag(score_tag) <= len(read.seq) - score_diff_max:
        return None
    if not read.has_tag(umi_tag):
        return None
    if read.mapq < mapq_threshold:
        return None
    return p_misaligned_default, hash_string(read.get_tag(umi_tag))


class SnpPositions:
    """Sorted SNP positions of one chromosome with the two queries the counting loop needs."""

    def __init__(self, positions: np.ndarray):
        positions = np.asarray(positions)
        assert np.array_equal(positions, np.sort(positions))
        self.positions = positions
        self._list = positions.tolist()

    def any_in(self, start: int, end: int) -> bool:
        """Is there a SNP in [start, end)?  (snp_counter.py:31-36)"""
        lo = int(np.searchsorted(self.positions, start, side='left'))
        return lo < len(self._list) and self._list[lo] < end

    def calls_of_read(self, read) -> List[Tuple[int, str, int]]:
        """(reference position, base, base quality) for every SNP an aligned block of the read covers
        (snp_counter.py:38-69; clips and insertions advance the read cursor, deletions / skips the reference)."""
        out: List[Tuple[int, str, int]] = []
        if not self.any_in(read.reference_start, read.reference_end + 1):
            return out
        seq, qual = read.seq, read.query_qualities
        in_read, in_ref = 0, read.reference_start
        for op, length in read.cigartuples:
            if op in (0, 7, 8):
                lo, hi = np.searchsorted(self.positions, [in_ref, in_ref + length])
                for position in self._list[lo:hi]:
                    k = in_read + (position - in_ref)
                    out.append((position, seq[k], qual[k]))
                in_ref += length
                in_read += length
            elif op in (2, 3):
                in_ref += length
            elif op in (1, 4, 5, 6):
                in_read += length
            else:
                raise NotImplementedError(f'cigar code unknown {op}')
        return out


def collapse_molecule(reads: List[Tuple[object, float]], snps: SnpPositions, skip_complete_duplicates: bool = True):
    """
    One (barcode, UMI) group -> (p_group_misaligned, [(position, base, p_base_wrong)]); snp_counter.py:142-192.
    Reads with identical (start, end, AS) count once; per position every observed base accumulates the product of
    0.1 ** (0.1 * min(q, 40)); candidates 1000x worse than the best are dropped; positions that still have more than
    one candidate base yield no call.
    """
    p_group = 1
    seen = set()
    per_position: Dict[int, List[Tuple[str, int]]] = {}
    for read, p_read in reads:
        if skip_complete_duplicates:
            signature = (read.reference_start, read.reference_end, read.get_tag('AS'))
            if signature in seen:
                continue
            seen.add(signature)
        p_group *= p_read
        for position, base, quality in snps.calls_of_read(read):
            per_position.setdefault(position, []).append((base, quality))
    calls = []
    for position, observations in per_position.items():
        wrong: Dict[str, float] = {}
        for base, quality in observations:
            wrong[base] = wrong.get(base, 1) * 0.1 ** (0.1 * min(quality, 40))
        if len(wrong) > 1:
            best = min(wrong.values())
            wrong = {base: p for base, p in wrong.items() if p <= best * 1000}
        if len(wrong) == 1:
            (base, p), = wrong.items()
            calls.append((position, base, p))
    return p_group, calls


# ---------------------------------------------------------------------------------------------------------------
# native path (csrc_host/bam_counter.cpp through the C ABI of include/demux_io.h)
# ---------------------------------------------------------------------------------------------------------------

_NATIVE_FILTERS = {}  # read-filter callback -> its parameters for the native loop


def _register_native_filter(fn, **params):
    _NATIVE_FILTERS[fn] = params


_register_native_filter(parse_read, umi_tag='UB', nhits_tag='NH', score_tag='AS', score_diff_max=8,
                        mapq_threshold=20, p_misaligned_default=0.01)
_register_native_filter(parse_read_bd_rhapsody, umi_tag='MA', nhits_tag='', score_tag='AS', score_diff_max=8,
                        mapq_threshold=20, p_misaligned_default=0.01)

_io_lib = None


def native_io():
    """ctypes handle of libdemux_io.so, or None when it has not been built (the Python loop is used then)."""
    global _io_lib
    if _io_lib is None:
        import ctypes as C
        from .build import HOST_LIB_PATH
        if not HOST_LIB_PATH.exists():
            _io_lib = False
        else:
            lib = C.CDLL(str(HOST_LIB_PATH))
            lib.dmxio_last_error.restype = C.c_char_p
            lib.dmxio_count_region.restype = C.c_void_p
            lib.dmxio_count_region.argtypes = [C.c_char_p, C.c_int32, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p,
                                               C.c_int64, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p,
                                               C.c_int32, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32,
                                               C.c_double]
            for name in ('dmxio_n_molecules', 'dmxio_n_calls', 'dmxio_n_reads_seen'):
                getattr(lib, name).restype = C.c_int64
                getattr(lib, name).argtypes = [C.c_void_p]
            lib.dmxio_count_coverage.restype = C.c_int
            lib.dmxio_count_coverage.argtypes = [C.c_char_p, C.c_int32, C.c_uint64, C.c_int64, C.c_int64, C.c_char_p,
                                                 C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
            lib.dmxio_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            lib.dmxio_copy.restype = None
            lib.dmxio_free.argtypes = [C.c_void_p]
            lib.dmxio_free.restype = None
            _io_lib = lib
    return _io_lib or None


_INFO_CACHE: Dict[str, dict] = {}


def _bam_info(path) -> dict:
    """Header (+ .bai index when present) of a BAM, without inflating the alignments."""
    key = str(path)
    if key not in _INFO_CACHE:
        names, lengths, first = read_bam_header(key)
        bai = None
        for candidate in (key + '.bai', key[:-4] + '.bai' if key.endswith('.bam') else None):
            if candidate and Path(candidate).exists():
                bai = read_bai(candidate)
                break
        _INFO_CACHE[key] = dict(names=names, lengths=lengths, first_voffset=first, bai=bai)
    return _INFO_CACHE[key]


def _whitelist_blob(barcode_handler: BarcodeHandler):
    """Whitelist keys as one byte blob + offsets + compressed ids (tuple keys (CB, RG) are joined with 0x1f).
    Built once per handler: every region task of a BAM passes the same whitelist."""
    mapping = barcode_handler.barcode2index
    stamp = (id(mapping), len(mapping))
    cached = getattr(barcode_handler, '_native_whitelist', None)
    if cached is not None and cached[0] == stamp:
        return cached[1]
    blob = _build_whitelist_blob(barcode_handler)
    try:
        barcode_handler._native_whitelist = (stamp, blob)
    except AttributeError:  # a handler type that does not take attributes: rebuild per task
        pass
    return blob


def _build_whitelist_blob(barcode_handler: BarcodeHandler):
    keys, ids = [], []
    for key, idx in barcode_handler.barcode2index.items():
        if isinstance(key, tuple):
            key = key[0] + '\x1f' + key[1]
        if isinstance(key, str):  # filter_to_rg_value() leaves integer placeholders for foreign barcodes
            keys.append(key.encode())
            ids.append(idx)
    offsets = np.zeros(len(keys) + 1, dtype=np.int64)
    np.cumsum([len(k) for k in keys], out=offsets[1:])
    return b''.join(keys), offsets, np.asarray(ids, dtype=np.int32)


def _count_region_native(lib, path: str, chromosome: str, positions: np.ndarray, barcode_handler: BarcodeHandler,
                         params: dict, start, stop) -> CompressedSNPCalls:
    info = _bam_info(path)
    ref_id = info['names'].index(chromosome)
    voffset = info['first_voffset']
    if info['bai'] is not None:
        from_index = region_start_voffset(info['bai'][ref_id], start)
        voffset = from_index if from_index is not None else voffset
    positions = np.ascontiguousarray(positions, dtype=np.int64)
    blob, offsets, ids = _whitelist_blob(barcode_handler)
    handle = lib.dmxio_count_region(
        str(path).encode(), ref_id, voffset, -1 if start is None else int(start), -1 if stop is None else int(stop),
        positions.ctypes.data, len(positions), blob, offsets.ctypes.data, ids.ctypes.data, len(ids),
        barcode_handler.tag.encode(), int(barcode_handler.use_rg), params['umi_tag'].encode(),
        params['nhits_tag'].encode(), params['score_tag'].encode(), params['score_diff_max'],
        params['mapq_threshold'], params['p_misaligned_default'])
    if not handle:
        raise RuntimeError(f'native count_snps failed: {lib.dmxio_last_error().decode()}')
    try:
        out = CompressedSNPCalls.__new__(CompressedSNPCalls)
        out.n_molecules, out.n_snp_calls = int(lib.dmxio_n_molecules(handle)), int(lib.dmxio_n_calls(handle))
        out.molecules = np.empty(out.n_molecules, dtype=MOLECULE_DTYPE)
        out.snp_calls = np.empty(out.n_snp_calls, dtype=SNP_CALL_DTYPE)
        lib.dmxio_copy(handle, out.molecules.ctypes.data, out.snp_calls.ctypes.data)
    finally:
        lib.dmxio_free(handle)
    return out


def count_coverage_native(path, chromosome: str, start: int, stop: int, parse_read: Callable,
                          quality_threshold: int = 15) -> Optional[np.ndarray]:
    """int32 [4, stop - start] coverage of reads passing a built-in read filter, by the native loop; None when the
    filter is a custom callback or libdemux_io.so is not built (the caller then walks the reads in Python)."""
    lib, params = native_io(), _NATIVE_FILTERS.get(parse_read)
    if lib is None or params is None:
        return None
    info = _bam_info(path)
    ref_id = info['names'].index(chromosome)
    voffset = info['first_voffset']
    if info['bai'] is not None:
        from_index = region_start_voffset(info['bai'][ref_id], start)
        voffset = from_index if from_index is not None else voffset
    counts = np.zeros((4, max(int(stop) - int(start), 0)), dtype=np.int32)
    if counts.size:
        rc = lib.dmxio_count_coverage(str(path).encode(), ref_id, voffset, int(start), int(stop),
                                      params['umi_tag'].encode(), params['nhits_tag'].encode(),
                                      params['score_tag'].encode(), params['score_diff_max'], params['mapq_threshold'],
                                      int(quality_threshold), counts.ctypes.data)
        if rc != 0:
            raise RuntimeError(f'native count_coverage failed: {lib.dmxio_last_error().decode()}')
    return counts


def count_region(bamfile, chromosome: str, positions: np.ndarray, barcode_handler: BarcodeHandler,
                 parse_read: Callable, start=None, stop=None, use_native: bool = True) -> Tuple[str, CompressedSNPCalls]:
    """One counting task (snp_counter.py:234-276): native loop for the built-in read filters, Python otherwise."""
    lib = native_io() if use_native else None
    if lib is not None and parse_read in _NATIVE_FILTERS and not isinstance(bamfile, BamFile):
        return chromosome, _count_region_native(lib, str(bamfile), chromosome, positions, barcode_handler,
                                                _NATIVE_FILTERS[parse_read], start, stop)
    bam = bamfile if isinstance(bamfile, BamFile) else _open_cached(bamfile)
    snps = SnpPositions(positions)
    out = CompressedSNPCalls()
    open_groups: Dict[Tuple[int, int], list] = {}  # (cb, ub) -> [furthest reference_end, [(read, p_misaligned)]]

    def flush(threshold) -> None:
        closed = [key for key, (reach, _reads) in open_groups.items() if reach < threshold]
        for key in closed:
            _reach, reads = open_groups.pop(key)
            if not snps.any_in(min(r.reference_start for r, _ in reads), max(r.reference_end for r, _ in reads) + 1):
                continue
            p_group, calls = collapse_molecule(reads, snps)
            if calls:
                out.add_calls_from_read_group(key[0], key[1], p_group, calls)

    previous_segment = None
    for read in bam.fetch(chromosome, start=start, stop=stop):
        parsed = parse_read(read)
        if parsed is None:
            continue
        cb = barcode_handler.get_barcode_index(read)
        if cb is None:
            continue
        p_misaligned, ub = parsed
        group = open_groups.get((cb, ub))
        if group is None:
            open_groups[(cb, ub)] = [read.reference_end, [(read, p_misaligned)]]
        else:
            group[0] = max(group[0], read.reference_end)
            group[1].append((read, p_misaligned))
        segment = read.reference_start // SEGMENT_LENGTH
        if segment != previous_segment:
            flush(read.reference_start - SEGMENT_LENGTH)
            previous_segment = segment
    flush(float('inf'))
    out.minimize_memory_footprint()
    return chromosome, out


_BAM_CACHE: Dict[str, BamFile] = {}


def _open_cached(path) -> BamFile:
    key = str(path)
    if key not in _BAM_CACHE:
        _BAM_CACHE.clear()  # one inflated BAM per process is enough
        _BAM_CACHE[key] = BamFile(key)
    return _BAM_CACHE[key]


def plan_tasks(bamfile_location, chromosome2positions: Dict[str, np.ndarray], barcode_handler: BarcodeHandler,
               n_reads_per_job: int = 10_000_000, minimum_fragment_length_per_job: int = 5_000,
               minimum_overlap: int = 100) -> list:
    """Region tasks, most complex first (snp_counter.py:330-385); a dict of BAMs keyed by RG fans out per file."""
    if isinstance(bamfile_location, dict):
        assert barcode_handler.use_rg, 'barcode handler should use RG tag'
        tasks = []
        for rg in set(rg for _barcode, rg in barcode_handler.barcode2index):
            assert rg in bamfile_location, f'{rg} has no matching path in bamfile_location parameter'
            tasks.extend(plan_tasks(bamfile_location[rg], chromosome2positions, barcode_handler.filter_to_rg_value(rg),
                                    n_reads_per_job, minimum_fragment_length_per_job, minimum_overlap))
        return tasks
    info = _bam_info(bamfile_location)
    if info['bai'] is not None:  # what pysam's get_index_statistics() reads
        mapped = {name: entry['mapped'] for name, entry in zip(info['names'], info['bai'])}
    else:
        mapped = _open_cached(bamfile_location).mapped_reads_per_reference()
    ranked = []
    for chromosome, positions in chromosome2positions.items():
        length = info['lengths'][info['names'].index(chromosome)]
        n_jobs = max(1, min(mapped[chromosome] // n_reads_per_job, length // minimum_fragment_length_per_job))
        cuts = np.searchsorted(positions, np.linspace(0, length, n_jobs + 1)[1:-1])
        for subset in np.split(positions, cuts):
            if len(subset) == 0:
                continue
            start = max(0, min(subset) - minimum_overlap)
            stop = min(length, max(subset) + minimum_overlap)
            complexity = len(subset) * mapped[chromosome] / length ** 0.5
            ranked.append((complexity, (bamfile_location, chromosome, start, stop, subset, barcode_handler)))
    return [task for _complexity, task in sorted(ranked, reverse=True)]


def _resolve_n_jobs(n_jobs) -> int:
    """joblib's convention: -1 = all cores, -2 = all but one, ..."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    if n_jobs is None:
        return 1
    n_jobs = int(n_jobs)
    return max(1, cores + 1 + n_jobs) if n_jobs < 0 else max(1, n_jobs)


def count_snps(bamfile_location, chromosome2positions: Dict[str, np.ndarray], barcode_handler: BarcodeHandler,
               joblib_n_jobs=-1, joblib_verbosity=11, parse_read=parse_read,
               use_native: bool = True) -> Dict[str, CompressedSNPCalls]:
    """
    Which molecules carry information about which SNPs: {chromosome: CompressedSNPCalls}, the input of
    `Demultiplexer.predict_posteriors / learn_genotypes`.  Arguments as in the reference (snp_counter.py:279-302);
    `bamfile_location` may be a path or, with an RG-aware barcode handler, a dict RG -> path.  With the built-in
    read filters (`parse_read`, `parse_read_bd_rhapsody`) the per-read loop runs in native code (libdemux_io.so)
    with identical output; any other callback, or `use_native=False`, takes the Python loop.
    """
    tasks = plan_tasks(bamfile_location, chromosome2positions, barcode_handler)

    def run(task):
        bamfile, chromosome, start, stop, positions, handler = task
        return count_region(bamfile, chromosome, positions, handler, parse_read, start=start, stop=stop,
                            use_native=use_native)

    native = use_native and native_io() is not None and parse_read in _NATIVE_FILTERS
    if joblib_n_jobs == 1 or len(tasks) <= 1:
        results = [run(task) for task in tasks]
    elif native:
        # the native loop runs outside the GIL (ctypes): threads of this process, nothing to pickle back (the
        # reference's worker processes return tens of MB per task); same tasks, same order of results
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(_resolve_n_jobs(joblib_n_jobs), len(tasks))) as pool:
            results = list(pool.map(run, tasks))
    else:
        import joblib
        with joblib.Parallel(n_jobs=joblib_n_jobs, verbose=joblib_verbosity, pre_dispatch='all') as parallel:
            results = parallel(joblib.delayed(count_region)(bamfile, chromosome, positions, handler, parse_read,
                                                            start=start, stop=stop, use_native=use_native)
                               for bamfile, chromosome, start, stop, positions, handler in tasks)
    per_chromosome = defaultdict(list)
    for chromosome, calls in results:
        per_chromosome[chromosome].append(calls)
    # a chromosome counted by one task keeps that task's (tight, unshared) arrays: nothing to merge or copy
    return {chromosome: parts[0] if len(parts) == 1 else CompressedSNPCalls.concatenate(parts)
            for chromosome, parts in per_chromosome.items()}
