"""
`Demultiplexer` -- drop-in for demuxalot.Demultiplexer (demuxalot/demux.py:24-392) whose numeric stages
run as hand-written sm_100a CUDA kernels behind the C ABI of include/demux_b200.h.

Host side (this file) is Python, as the reference is: it flattens the inputs, owns the device buffers
(torch tensors used purely as memory + streams), calls the kernels through ctypes and assembles the
DataFrames / genotype objects the reference returns.  No numeric stage has a CPU implementation here; without
a CUDA device or without libdemux_b200.so every entry point raises.

Stage map (reference lines -> ABI call):
  pack_calls                     demux.py:302-392  -> dmx_unpack_match_calls, dmx_build_rows, dmx_prior_betas
  _compute_probs_from_betas      demux.py:267-274  -> dmx_probs_from_betas
  compute_barcode_logits + softmax  demux.py:246-265,101,152 -> dmx_estep
  M-step loop                    demux.py:113-118  -> dmx_mstep
  compute_barcode_logits, aggregate_on_snps=True  demux.py:204-244 -> dmx_build_snp_groups, dmx_snp_logits,
                                                                      dmx_softmax_rows_f64
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import threading
import warnings
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from . import _native
from .calls import MOLECULE_DTYPE, SNP_CALL_DTYPE


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError('demuxalot_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback')


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _to_device(array: np.ndarray, device) -> torch.Tensor:
    """Host numpy -> device tensor.  Structured arrays travel as their raw packed bytes."""
    array = np.ascontiguousarray(array)
    if array.dtype.fields is not None:
        array = array.view(np.uint8)
    with warnings.catch_warnings():
        # read-only inputs (e.g. genotypes.get_betas()) are only ever read through this tensor
        warnings.filterwarnings('ignore', message='The given NumPy array is not writable')
        host = torch.from_numpy(array)
    return host.to(device, non_blocking=True)


_COPY_STREAMS: Dict[int, 'torch.cuda.Stream'] = {}


def _copy_stream(device) -> 'torch.cuda.Stream':
    """One side stream per device for host->device uploads that overlap the kernels of the current stream."""
    key = torch.device(device).index
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


def _upload_async(array: np.ndarray, device, copy_stream):
    """Starts the upload on `copy_stream`; returns (tensor, event).  The consumer stream must wait for the event."""
    consumer = torch.cuda.current_stream()
    with torch.cuda.stream(copy_stream):
        tensor = _to_device(array, device)
        event = torch.cuda.Event()
        event.record(copy_stream)
    tensor.record_stream(consumer)  # allocated on the copy stream, read by kernels of the consumer stream
    return tensor, event


_CB_STAGING: Dict[int, dict] = {}
_CB_STAGING_LOCK = threading.Lock()  # one packing call at a time fills and uploads the per-device staging buffer


def _cb_staging(device, n_int32: int) -> dict:
    """Pinned host buffer (per device, grown on demand) through which the compressed_cb column travels, with the
    event of the last upload that read it."""
    key = torch.device(device).index
    entry = _CB_STAGING.get(key)
    if entry is None or entry['buffer'].numel() < n_int32:
        if entry is not None and entry['last_read'] is not None:
            entry['last_read'].synchronize()
        entry = {'buffer': torch.empty(max(n_int32, 1), dtype=torch.int32, pin_memory=True), 'last_read': None}
        _CB_STAGING[key] = entry
    return entry


def _device_index(index: dict, device) -> dict:
    """Device copy of the genotype index (sorted keys, SNP CSR), cached with the host index it was made from."""
    cache = index.setdefault('_device', {})
    key = torch.device(device).index
    if key not in cache:
        cache[key] = {name: _to_device(index[name], device)
                      for name in ('keys_sorted', 'vids_sorted', 'snp_offsets', 'snp_variants')}
        torch.cuda.current_stream().synchronize()  # pageable source buffers: finish before anything reuses them
    return cache[key]


def _to_host(*tensors: torch.Tensor):
    """Device -> host through pinned staging buffers (torch caches them), one synchronisation for all."""
    staged = []
    for t in tensors:
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        staged.append(h)
    torch.cuda.current_stream().synchronize()
    return [h.numpy() for h in staged]


def n_options(n_genotypes: int, doublet_prior: float) -> int:
    return n_genotypes if doublet_prior == 0 else n_genotypes * (n_genotypes + 1) // 2


def option_names(genotype_names, doublet_prior: float) -> List[str]:
    """Column names in the order of demux.py:175-191: singlets, then 'Gi+Gj' for i < j, i-major."""
    names = list(genotype_names)
    out = list(names)
    if doublet_prior != 0:
        assert doublet_prior > 0
        for i, a in enumerate(names):
            out.extend(f'{a}+{b}' for b in names[i + 1:])
    return out


@dataclass
class DevicePack:
    """Device-resident result of pack_calls: rows in both orders, SNP index, regularised betas."""
    device: torch.device
    n_barcodes: int
    n_variants: int
    n_genotypes: int
    n_snps: int
    n_calls: int
    n_matched: int
    n_rows: int
    barcode_range: Tuple[int, int]  # global barcodes [lo, hi) held by this pack; ids below are local (cb - lo)
    # molecule-level calls in original order (variant == -1: unmatched); None when the caller does not need them
    call_variant: Optional[torch.Tensor]
    call_cb: Optional[torch.Tensor]
    call_e: Optional[torch.Tensor]
    # rows, reference order (variant-major) -- M-step
    csc_variant: torch.Tensor
    csc_cb: torch.Tensor
    csc_e: torch.Tensor
    csc_count: torch.Tensor
    variant_offsets: torch.Tensor
    # rows, barcode-major -- E-step
    csr_variant: torch.Tensor
    csr_e: torch.Tensor
    csr_row: torch.Tensor
    barcode_offsets: torch.Tensor
    barcode_order: torch.Tensor  # launch schedule: barcodes by descending row count
    n_mol: torch.Tensor
    snp_offsets: torch.Tensor
    snp_variants: torch.Tensor
    variant2snp: np.ndarray
    raw_betas: torch.Tensor
    betas: torch.Tensor  # regularised betas (demux.py:388), float32 [V, G]
    betas_min: Optional[torch.Tensor] = None  # min of the raw betas (device scalar), see check()

    def check(self, betas_min=None) -> None:
        """demux.py:374 -- negative betas are rejected.  Reads the device scalar unless the caller already
        downloaded it together with its results (one synchronisation instead of two)."""
        if betas_min is None and self.betas_min is not None:
            betas_min = float(self.betas_min)
        if betas_min is not None:
            assert float(betas_min) >= 0, 'bad genotypes provided, negative betas appeared'


class Demultiplexer:
    """
    Demultiplexer that can infer (learn) additional information about genotypes to achieve better quality.
    Same static API and class attributes as the reference (demux.py:24-32).
    """
    contribution_power = 2.
    # True selects the reference's experimental per-(barcode, SNP) regularised likelihood (demux.py:204-244): float64
    # logits / posteriors, no doublet penalties -- quirks of the reference kept as they are (_e_step_aggregated)
    aggregate_on_snps = False
    compensation_during_computing_barcode_logits = 0.5

    # B200-specific knobs (not in the reference)
    # 'auto' (reference roundings for doublet_prior == 0, product arithmetic with doublet columns) | 'fast' | 'exact',
    # see include/demux_b200.h DMX_ESTEP_*
    estep_flavour = 'auto'
    device: Optional[torch.device] = None  # None -> current CUDA device
    process_group = None  # torch.distributed group for barcode-sharded / multi-lane EM (see distributed.py)
    schedule_barcodes = True  # launch the deepest barcodes first (dmx_barcode_schedule)
    # warp-per-item pair E-step (dmx_estep_plan): barcodes deeper than this many rows are cut into segments
    # (16..4096); 0 disables the plan and every width runs on the CTA-per-barcode kernel
    estep_segment_rows = 4096
    planned_mstep = True  # three-tier M-step schedule (dmx_mstep_plan); False: one warp per variant for all
    pipelined_upload = True  # host->device copies on a side stream, overlapping the unpack / row-builder kernels
    # Only `compressed_cb` of the 12-byte molecule records is read (demux.py:352).  A pool of host threads copies that
    # column into a pinned staging buffer while the snp_calls records are on the wire, and 4 instead of 12 bytes per
    # molecule are uploaded.  0 threads: upload the records as they are.
    # (default: half of the host cores, shared between the ranks of the node, at most 16)
    host_gather_threads = min(16, max(1, (os.cpu_count() or 2) // (2 * max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1'))))))
    # variant-range tiles of the sharded M-step: all-reduce of tile k overlaps the M-step of tile k + 1.  Measured
    # (profiles/r01_allreduce_sweep_*.json): the 168 MB all-reduce is 0.34 ms over NVLink, every extra tile costs
    # ~0.25 ms of stream hand-over, so one tile wins; more tiles only pay off for tables of many GB.
    pack_profile: Optional[dict] = None  # set to a dict to receive the wall time of the pack stages (synchronises)
    # How the M-step partials of the barcode shards are summed across GPUs:
    #   'peer': dmx_peer_sum_f32 -- our own reduce-scatter + all-gather kernel over NVLink peer memory (float32 partials
    #           summed in rank order in float64, one rounding; identical bits on every rank); needs torch symmetric memory
    #   'nccl': dmx_mstep_allreduce -- NCCL collectives on a second stream, tiled against the M-step kernel
    #   'auto': 'peer' when the platform offers it, else 'nccl'
    mstep_exchange = 'auto'
    last_exchange: Optional[str] = None  # diagnostic: the exchange the last sharded M-step actually ran
    # tiles of the NCCL path; 0 = by table size: 1 below 1 GiB (every tile costs a launch tail of the M-step tiers, which
    # outweighs what it hides: profiles/r02_sweep_mstep_allreduce_n2.json), 4 above
    mstep_allreduce_tiles = 0
    # 'float64': reduce-scatter of float64 partials, one rounding after the global sum (N GPUs give the bits of one up
    # to float64 regrouping), all-gather of float32; 'float32': float32 all-reduce, half the bytes again at the price of
    # one rounding per shard (~world_size * 6e-8 relative on the addition, far inside the 1e-5 parity bar)
    mstep_allreduce_dtype = 'float32'

    # ------------------------------------------------------------------------------------------------ helpers
    @classmethod
    def _device(cls) -> torch.device:
        _require_cuda()
        return torch.device('cuda', torch.cuda.current_device()) if cls.device is None else torch.device(cls.device)

    @classmethod
    def _flavour(cls) -> int:
        return {'fast': _native.ESTEP_FAST, 'exact': _native.ESTEP_EXACT, 'auto': _native.ESTEP_AUTO}[cls.estep_flavour]

    @staticmethod
    def _doublet_penalties(n_genotypes: int, doublet_prior: float) -> np.ndarray:
        """float32 [C] logit offsets (demux.py:158-173); host helper, the kernel applies the same constant."""
        assert 0 <= doublet_prior < 1
        out = np.zeros(n_options(n_genotypes, doublet_prior), dtype='float32')
        if doublet_prior != 0:
            bonus = np.log(n_genotypes * doublet_prior)
            bonus -= np.log(n_genotypes * max(n_genotypes - 1, 1) / 2 * (1 - doublet_prior))
            out[n_genotypes:] = bonus
        return out

    # ------------------------------------------------------------------------------------------------ pack
    @staticmethod
    def _genotype_index(genotypes) -> dict:
        return genotypes.hot_path_index() if hasattr(genotypes, 'hot_path_index') else _foreign_hot_index(genotypes)

    @staticmethod
    def _select_parts(chromosome2compressed_snp_calls, chrom2id) -> list:
        """[(chrom_id, calls)] of the chromosomes that carry calls, in dict order (demux.py:334)."""
        parts = []
        for chrom, calls in chromosome2compressed_snp_calls.items():
            if chrom not in chrom2id:
                # demux.py:339-341 skips the chromosome and the counter check at :359 then fails
                assert calls.n_snp_calls == 0, \
                    f'calls on chromosome {chrom!r}, which is absent from the genotypes'
                continue
            if calls.n_snp_calls:
                parts.append((chrom2id[chrom], calls))
        return parts

    @classmethod
    def _upload_unpack(cls, parts, dindex, n_variants: int, dev, shard=None):
        """
        (a2) Uploads the packed records and matches them against the genotype keys (demux.py:334-358):
        returns (call_variant, call_cb, call_e, n_calls), device SoA over the calls this process handles, in call
        order.  `shard` = (rank, world, group): this rank takes the rank-th contiguous slice of every chromosome's
        calls (1 / world of the bytes cross its PCIe link); the compressed_cb column of the molecules is uploaded in
        slices as well and all-gathered over NVLink, since a call may point at any molecule.
        """
        lib = _native.load()
        stream = _stream()
        main = torch.cuda.current_stream()
        copy_stream = _copy_stream(dev) if cls.pipelined_upload else main
        gkeys, gvids = dindex['keys_sorted'], dindex['vids_sorted']
        rank, world, group = shard if shard is not None else (0, 1, None)

        def call_slice(n):
            return (n * rank // world, n * (rank + 1) // world)

        def molecule_slice(n):  # equal slices (all_gather_into_tensor), the last ones possibly short or empty
            per = -(-n // world) if n else 0
            return min(rank * per, n), min((rank + 1) * per, n), per

        call_ranges = [call_slice(c.n_snp_calls) for _cid, c in parts]
        n_calls = sum(hi - lo for lo, hi in call_ranges)
        call_variant = torch.empty(max(n_calls, 1), dtype=torch.int32, device=dev)
        call_cb = torch.empty(max(n_calls, 1), dtype=torch.int32, device=dev)
        call_e = torch.empty(max(n_calls, 1), dtype=torch.float32, device=dev)
        # All uploads are queued on the copy stream up front; the unpack kernel of chromosome k runs on the main
        # stream while chromosome k + 1 is still on the wire.
        gather = (cls.host_gather_threads > 0 or world > 1) and len(parts) > 0
        staging = ready_flags = worker = None
        molecule_arrays = [_as_dtype(c.molecules[:c.n_molecules], MOLECULE_DTYPE) for _cid, c in parts]
        mol_ranges = [molecule_slice(len(m)) for m in molecule_arrays]
        # one packing call at a time fills the per-device staging buffer and queues the uploads out of it
        with (_CB_STAGING_LOCK if gather else contextlib.nullcontext()):
            if gather:
                staging = _cb_staging(dev, sum(hi - lo for lo, hi, _per in mol_ranges))
                if staging['last_read'] is not None:
                    staging['last_read'].synchronize()  # the previous call's upload out of this buffer is done
                ready_flags = [threading.Event() for _ in parts]
                base_ptr, n_threads = staging['buffer'].data_ptr(), max(1, int(cls.host_gather_threads))

                def gather_all():  # ctypes releases the GIL: runs beside the upload submissions below
                    offset = 0
                    for k, (mols, (lo, hi, _per)) in enumerate(zip(molecule_arrays, mol_ranges)):
                        rc = -1  # whatever happens the flag is set, so the consumer below never waits for ever
                        try:
                            rc = lib.dmx_host_gather_cb(mols[lo:hi].ctypes.data, hi - lo, base_ptr + 4 * offset, n_threads)
                        finally:
                            # dmx_last_error() is thread-local: read it on THIS thread, the consumer runs on another
                            ready_flags[k].msg = lib.dmx_last_error().decode(errors='replace') if rc else ''
                            ready_flags[k].rc = rc
                            ready_flags[k].set()
                        offset += hi - lo

                worker = threading.Thread(target=gather_all, daemon=True)
                worker.start()
            uploads = []
            for (cid, calls), molecules, (k_lo, k_hi) in zip(parts, molecule_arrays, call_ranges):
                snp_calls = _as_dtype(calls.snp_calls[:calls.n_snp_calls], SNP_CALL_DTYPE)[k_lo:k_hi]
                d_calls, ready = _upload_async(snp_calls, dev, copy_stream)
                d_mols = None
                if not gather:
                    d_mols, ready = _upload_async(molecules, dev, copy_stream)
                uploads.append([cid, k_hi - k_lo, d_calls, d_mols, len(molecules), ready])
            if gather:  # the compact columns follow the call records on the wire, in the order the gathers finish
                offset = 0
                for k, entry in enumerate(uploads):
                    ready_flags[k].wait()
                    if ready_flags[k].rc:
                        raise _native.NativeError(f'dmx_host_gather_cb failed (rc={ready_flags[k].rc}): {ready_flags[k].msg}')
                    lo, hi, per = mol_ranges[k]
                    with torch.cuda.stream(copy_stream):
                        if world > 1:  # equal-sized slices: the tail of the last ones is never read
                            d_cb = torch.zeros(max(per, 1), dtype=torch.int32, device=dev)
                            d_cb[:hi - lo].copy_(staging['buffer'][offset:offset + hi - lo], non_blocking=True)
                        else:
                            d_cb = staging['buffer'][offset:offset + hi - lo].to(dev, non_blocking=True)
                        event = torch.cuda.Event()
                        event.record(copy_stream)
                    d_cb.record_stream(main)
                    entry[3], entry[5] = d_cb, event
                    staging['last_read'] = event
                    offset += hi - lo
                worker.join()
        done = 0
        for (cid, n, d_calls, d_mols, n_molecules, ready), (_lo, _hi, per) in zip(uploads, mol_ranges):
            main.wait_event(ready)
            if world > 1:
                import torch.distributed as dist
                full = torch.empty(max(per, 1) * world, dtype=torch.int32, device=dev)
                dist.all_gather_into_tensor(full, d_mols, group=group)
                d_mols = full
            if n:
                _native.check(lib.dmx_unpack_match_calls(
                    d_calls.data_ptr(), n, d_mols.data_ptr(), n_molecules, 4 if gather else 12, cid,
                    gkeys.data_ptr(), gvids.data_ptr(), n_variants,
                    call_variant[done:].data_ptr(), call_cb[done:].data_ptr(), call_e[done:].data_ptr(), stream),
                    'dmx_unpack_match_calls')
            done += n
        del uploads  # the record buffers return to the allocator once the kernels that read them are done
        return call_variant, call_cb, call_e, n_calls

    @classmethod
    def _unpack_device_parts(cls, device_parts, dindex, chrom2id, n_variants: int, dev):
        """(a2) for inputs that already sit in device memory: device_parts = [dict(chromosome, records uint8 [n, 13],
        molecule_cb int32 [n_molecules], n_calls)] (what the uploads of _upload_unpack leave behind)."""
        lib = _native.load()
        n_calls = sum(int(p['n_calls']) for p in device_parts)
        call_variant = torch.empty(max(n_calls, 1), dtype=torch.int32, device=dev)
        call_cb = torch.empty(max(n_calls, 1), dtype=torch.int32, device=dev)
        call_e = torch.empty(max(n_calls, 1), dtype=torch.float32, device=dev)
        done = 0
        for part in device_parts:
            n = int(part['n_calls'])
            if n:
                _native.check(lib.dmx_unpack_match_calls(
                    part['records'].data_ptr(), n, part['molecule_cb'].data_ptr(), part['molecule_cb'].numel(), 4,
                    chrom2id[part['chromosome']], dindex['keys_sorted'].data_ptr(), dindex['vids_sorted'].data_ptr(),
                    n_variants, call_variant[done:].data_ptr(), call_cb[done:].data_ptr(), call_e[done:].data_ptr(),
                    _stream()), 'dmx_unpack_match_calls')
            done += n
        return call_variant, call_cb, call_e, n_calls

    @classmethod
    def _route_calls(cls, call_variant, call_cb, call_e, n_calls: int, n_variants: int, n_barcodes: int, shard, dev):
        """
        Sharded pack, step 2: contiguous barcode ranges balanced by matched calls (histogram all-reduced over the
        ranks), stable partition of this rank's matched calls by owner (dmx_route_calls) and one all-to-all over
        NVLink.  Returns (call_variant, call_cb LOCAL to the range, call_e, n_calls, (lo, hi)) of this rank's shard,
        call order kept inside every chromosome.
        """
        import torch.distributed as dist
        from .distributed import exchange_calls, plan_barcode_shards
        lib = _native.load()
        rank, world, group = shard
        histogram = torch.zeros(max(n_barcodes, 1), dtype=torch.int64, device=dev)
        _native.check(lib.dmx_barcode_histogram(call_variant.data_ptr(), call_cb.data_ptr(), n_calls, n_variants,
                                                n_barcodes, histogram.data_ptr(), _stream()), 'dmx_barcode_histogram')
        dist.all_reduce(histogram, op=dist.ReduceOp.SUM, group=group)
        ranges = plan_barcode_shards(histogram[:n_barcodes].cpu().numpy(), world)  # identical on every rank
        cuts = torch.tensor([r[0] for r in ranges] + [n_barcodes], dtype=torch.int64, device=dev)
        ws_bytes = lib.dmx_route_calls_workspace_bytes(n_calls)
        if ws_bytes < 0:
            _native.check(-1, 'dmx_route_calls_workspace_bytes')
        workspace = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        send = [torch.empty(max(n_calls, 1), dtype=dt, device=dev) for dt in (torch.int32, torch.int32, torch.float32)]
        h_counts = (C.c_int64 * (world + 1))()
        _native.check(lib.dmx_route_calls(
            call_variant.data_ptr(), call_cb.data_ptr(), call_e.data_ptr(), n_calls, n_variants, n_barcodes,
            cuts.data_ptr(), world, workspace.data_ptr(), ws_bytes, send[0].data_ptr(), send[1].data_ptr(),
            send[2].data_ptr(), h_counts, _stream()), 'dmx_route_calls')
        del workspace
        counts = [int(c) for c in h_counts]
        received, n_received, n_bad = exchange_calls(send, counts[:world], counts[world], group, dev)
        # the reference would index past its per-barcode arrays: an input error, raised on every rank
        assert n_bad == 0, f'{n_bad} matched calls carry a compressed_cb outside [0, n_barcodes={n_barcodes})'
        return received[0], received[1], received[2], n_received, ranges[rank]

    @classmethod
    def _keep_barcode_range(cls, call_variant, call_cb, call_e, n_calls: int, n_variants: int, n_barcodes: int,
                            barcode_range: Tuple[int, int], dev):
        """Single-process restriction to the barcodes [lo, hi): the routing step of the sharded pack with three
        destinations (below / inside / above the range), of which the middle block is kept -- matched calls of the
        range in call order, barcode ids local to it."""
        lib = _native.load()
        lo, hi = barcode_range
        assert 0 <= lo <= hi <= n_barcodes
        cuts = torch.tensor([0, lo, hi, n_barcodes], dtype=torch.int64, device=dev)
        ws_bytes = lib.dmx_route_calls_workspace_bytes(n_calls)
        workspace = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        out = [torch.empty(max(n_calls, 1), dtype=dt, device=dev) for dt in (torch.int32, torch.int32, torch.float32)]
        h_counts = (C.c_int64 * 4)()
        _native.check(lib.dmx_route_calls(
            call_variant.data_ptr(), call_cb.data_ptr(), call_e.data_ptr(), n_calls, n_variants, n_barcodes,
            cuts.data_ptr(), 3, workspace.data_ptr(), ws_bytes, out[0].data_ptr(), out[1].data_ptr(),
            out[2].data_ptr(), h_counts, _stream()), 'dmx_route_calls')
        assert h_counts[3] == 0, f'{h_counts[3]} matched calls carry a compressed_cb outside [0, n_barcodes={n_barcodes})'
        first, n = int(h_counts[0]), int(h_counts[1])
        return tuple(t[first:first + max(n, 1)] for t in out) + (n,)

    @classmethod
    def _finish_pack(cls, call_variant, call_cb, call_e, n_calls: int, genotypes, index, dindex, n_barcodes: int,
                     add_data_prior: bool, dev, barcode_range: Tuple[int, int]) -> DevicePack:
        """(a3) + (a4): rows in both orders, schedule, regularised betas (demux.py:276-300, 362-390)."""
        lib = _native.load()
        stream = _stream()
        main = torch.cuda.current_stream()
        n_variants, n_genotypes = genotypes.n_variants, genotypes.n_genotypes
        raw = np.asarray(genotypes.get_betas())
        assert raw.dtype == np.float32 and raw.shape == (n_variants, n_genotypes)
        snp_offsets, snp_variants = dindex['snp_offsets'], dindex['snp_variants']
        if raw.size:
            raw_dev, raw_ready = _upload_async(raw, dev, _copy_stream(dev) if cls.pipelined_upload else main)
        else:
            raw_dev, raw_ready = torch.empty((n_variants, n_genotypes), device=dev), None

        ws_bytes = lib.dmx_build_rows_workspace_bytes(n_calls, n_variants, n_barcodes)
        if ws_bytes < 0:
            _native.check(-1, 'dmx_build_rows_workspace_bytes')
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        cap = max(n_calls, 1)
        i32 = dict(dtype=torch.int32, device=dev)
        csc_variant, csc_cb, csc_count = (torch.empty(cap, **i32) for _ in range(3))
        csr_variant, csr_row = (torch.empty(cap, **i32) for _ in range(2))
        csc_e = torch.empty(cap, dtype=torch.float32, device=dev)
        csr_e = torch.empty(cap, dtype=torch.float32, device=dev)
        variant_offsets = torch.empty(n_variants + 1, dtype=torch.int64, device=dev)
        barcode_offsets = torch.empty(n_barcodes + 1, dtype=torch.int64, device=dev)
        n_mol = torch.zeros(max(n_variants, 1), dtype=torch.int64, device=dev)  # filled only for the data prior
        h_rows, h_matched = C.c_int64(0), C.c_int64(0)
        _native.check(lib.dmx_build_rows(
            call_variant.data_ptr(), call_cb.data_ptr(), call_e.data_ptr(), n_calls, n_variants, n_barcodes,
            0, n_barcodes, workspace.data_ptr(), ws_bytes,
            csc_variant.data_ptr(), csc_cb.data_ptr(), csc_e.data_ptr(), csc_count.data_ptr(),
            variant_offsets.data_ptr(), csr_variant.data_ptr(), csr_e.data_ptr(), csr_row.data_ptr(),
            barcode_offsets.data_ptr(), n_mol.data_ptr() if add_data_prior else 0, C.byref(h_rows),
            C.byref(h_matched), stream),
            'dmx_build_rows')
        del workspace
        n_rows = int(h_rows.value)

        def fit(t):  # rows are usually ~60 % of the calls: give the rest of the allocation back
            return t[:n_rows].clone() if n_rows < 0.9 * cap and cap > (1 << 22) else t[:n_rows]

        csc_variant, csc_cb, csc_count, csc_e, csr_variant, csr_row, csr_e = (
            fit(t) for t in (csc_variant, csc_cb, csc_count, csc_e, csr_variant, csr_row, csr_e))
        barcode_order = torch.empty(max(n_barcodes, 1), dtype=torch.int32, device=dev)
        sched_bytes = lib.dmx_barcode_schedule_workspace_bytes(n_barcodes)
        sched_ws = torch.empty(max(sched_bytes, 1), dtype=torch.uint8, device=dev)
        _native.check(lib.dmx_barcode_schedule(barcode_offsets.data_ptr(), n_barcodes, barcode_order.data_ptr(),
                                               sched_ws.data_ptr(), sched_bytes, stream), 'dmx_barcode_schedule')

        if add_data_prior and cls.process_group is not None:
            # multi-lane / sharded EM: the data prior counts the molecules of every rank (demux.py:381)
            import torch.distributed as dist
            dist.all_reduce(n_mol, op=dist.ReduceOp.SUM, group=cls.process_group)

        if raw_ready is not None:
            main.wait_event(raw_ready)
        betas = torch.empty((n_variants, n_genotypes), dtype=torch.float32, device=dev)
        scratch = torch.empty(max(n_variants, 1), dtype=torch.float32, device=dev)
        _native.check(lib.dmx_prior_betas(
            raw_dev.data_ptr(), n_genotypes, n_variants, n_genotypes, snp_offsets.data_ptr(),
            snp_variants.data_ptr(), index['n_snps'], n_mol.data_ptr() if add_data_prior else 0,
            float(genotypes.default_prior), scratch.data_ptr(), betas.data_ptr(), n_genotypes, stream),
            'dmx_prior_betas')
        betas_min = raw_dev.min() if raw.size else None  # demux.py:374, asserted by DevicePack.check()

        return DevicePack(
            device=dev, n_barcodes=n_barcodes, n_variants=n_variants, n_genotypes=n_genotypes,
            n_snps=index['n_snps'], n_calls=n_calls, n_matched=int(h_matched.value), n_rows=n_rows,
            barcode_range=barcode_range,
            call_variant=call_variant[:n_calls], call_cb=call_cb[:n_calls], call_e=call_e[:n_calls],
            csc_variant=csc_variant, csc_cb=csc_cb, csc_e=csc_e, csc_count=csc_count, variant_offsets=variant_offsets,
            csr_variant=csr_variant, csr_e=csr_e, csr_row=csr_row,
            barcode_offsets=barcode_offsets, barcode_order=barcode_order[:n_barcodes], n_mol=n_mol[:n_variants],
            snp_offsets=snp_offsets, snp_variants=snp_variants, variant2snp=index['variant2snp'], raw_betas=raw_dev,
            betas=betas, betas_min=betas_min)

    @classmethod
    def _pack_device(cls, chromosome2compressed_snp_calls, genotypes, n_barcodes: int, add_data_prior: bool,
                     shard=None, device_parts=None, keep_calls: bool = True,
                     barcode_range: Optional[Tuple[int, int]] = None) -> DevicePack:
        """
        Device version of pack_calls (demux.py:302-392).
        shard = (rank, world, group): barcode-sharded pack -- this rank uploads a slice of the calls, the matched
        calls are exchanged, and the returned pack holds the barcodes [lo, hi) = pack.barcode_range with LOCAL
        barcode ids 0 .. hi - lo (pack.n_barcodes = hi - lo).
        device_parts: inputs already resident on the device (see _unpack_device_parts) instead of host records; with
        `shard` they are this rank's share of the calls (any subset).
        barcode_range = (lo, hi): single-process version of a shard -- only the barcodes [lo, hi) are kept, with local
        ids (what one rank of a sharded run holds, except that the data prior only counts this range's molecules).
        """
        dev = cls._device()
        index = cls._genotype_index(genotypes)
        n_variants = genotypes.n_variants
        profile = cls.pack_profile
        clock = [None]

        def lap(name):  # only when profiling: one synchronisation per stage
            if profile is not None:
                import time
                torch.cuda.synchronize(dev)
                now = time.perf_counter()
                if clock[0] is not None and name:
                    profile[name] = profile.get(name, 0.0) + now - clock[0]
                clock[0] = now

        with torch.cuda.device(dev):
            dindex = _device_index(index, dev)
            lap(None)
            if device_parts is not None:
                call_variant, call_cb, call_e, n_calls = cls._unpack_device_parts(
                    device_parts, dindex, index['chrom2id'], n_variants, dev)
            else:
                parts = cls._select_parts(chromosome2compressed_snp_calls, index['chrom2id'])
                call_variant, call_cb, call_e, n_calls = cls._upload_unpack(parts, dindex, n_variants, dev, shard)
            lap('unpack_match_s')
            if barcode_range is not None:
                assert shard is None
                call_variant, call_cb, call_e, n_calls = cls._keep_barcode_range(
                    call_variant, call_cb, call_e, n_calls, n_variants, n_barcodes, barcode_range, dev)
                n_barcodes = barcode_range[1] - barcode_range[0]
            else:
                barcode_range = (0, n_barcodes)
            if shard is not None:
                call_variant, call_cb, call_e, n_calls, barcode_range = cls._route_calls(
                    call_variant, call_cb, call_e, n_calls, n_variants, n_barcodes, shard, dev)
                n_barcodes = barcode_range[1] - barcode_range[0]
            lap('route_exchange_s')
            pack = cls._finish_pack(call_variant, call_cb, call_e, n_calls, genotypes, index, dindex, n_barcodes,
                                    add_data_prior, dev, barcode_range)
            lap('rows_and_prior_s')
            if not keep_calls:  # only pack_calls() and the aggregate_on_snps branch read the molecule-level calls
                pack.call_variant = pack.call_cb = pack.call_e = None
        return pack

    @classmethod
    def pack_calls(cls, chromosome2compressed_snp_calls, genotypes, add_data_prior: bool, n_barcodes: int = None):
        """
        Reference-shaped view of the packed calls (demux.py:302-392): returns
        (variant_index2snp_index, variant_index2betas, molecule_calls, barcode_calls) as host arrays.
        `barcode_calls` has the reference's fields except `barcode_snp_count`, which nothing downstream reads.
        The reference infers nothing about the number of barcodes here; pass `n_barcodes` or it is taken as
        max(compressed_cb) + 1.
        """
        if n_barcodes is None:
            n_barcodes = 1 + max((int(c.molecules['compressed_cb'][:c.n_molecules].max())
                                  for c in chromosome2compressed_snp_calls.values() if c.n_molecules), default=0)
        pack = cls._pack_device(chromosome2compressed_snp_calls, genotypes, n_barcodes, add_data_prior)
        pack.check()
        betas = pack.betas.cpu().numpy()
        betas.flags.writeable = False
        variant = pack.call_variant.cpu().numpy()
        keep = variant != -1
        molecule_calls = np.rec.fromarrays(
            [variant[keep], pack.variant2snp[variant[keep]], pack.call_cb.cpu().numpy()[keep],
             pack.call_e.cpu().numpy()[keep]],
            names=['variant_id', 'snp_id', 'compressed_cb', 'p_base_wrong'])
        rows_variant = pack.csc_variant.cpu().numpy()
        barcode_calls = np.rec.fromarrays(
            [rows_variant, pack.variant2snp[rows_variant], pack.csc_cb.cpu().numpy(), pack.csc_e.cpu().numpy(),
             pack.csc_count.cpu().numpy().astype(np.int64)],
            names=['variant_id', 'snp_id', 'compressed_cb', 'p_base_wrong', 'barcode_variant_count'])
        return pack.variant2snp.copy(), betas, molecule_calls, barcode_calls

    # ------------------------------------------------------------------------------------------------ stages
    @staticmethod
    def _table_ld(n_genotypes: int) -> int:
        return (n_genotypes + 3) // 4 * 4

    @classmethod
    def _probs_table(cls, pack: DevicePack, addition: Optional[torch.Tensor], p_genotype_clip: float,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(b) of north_star / demux.py:267-274 -> float32 [V, ld] table, ld = G rounded up to 4."""
        lib = _native.load()
        ld = cls._table_ld(pack.n_genotypes)
        if out is None:
            out = torch.empty((pack.n_variants, ld), dtype=torch.float32, device=pack.device)
        # probs.clip(p, 1 - p): bounds are Python floats cast to float32 (weak scalars), demux.py:274
        lo, hi = float(np.float32(p_genotype_clip)), float(np.float32(1 - p_genotype_clip))
        with torch.cuda.device(pack.device):
            _native.check(lib.dmx_probs_from_betas(
                pack.betas.data_ptr(), pack.n_genotypes, _native.ptr(addition), pack.n_genotypes, pack.n_variants,
                pack.n_genotypes, pack.snp_offsets.data_ptr(), pack.snp_variants.data_ptr(), pack.n_snps, lo, hi,
                out.data_ptr(), ld, _stream()), 'dmx_probs_from_betas')
        out.dmx_floor = lo  # every entry is >= the clip: lets the E-step pick its product length
        return out

    @classmethod
    def _estep_plan(cls, pack: DevicePack, doublet_prior: float):
        """Work items of the warp pair kernel for this pack: (seg_prefix, item_slot, n_items, seg_rows) or None.
        Computed once per pack and schedule setting (the plan only depends on the barcode depths)."""
        lib = _native.load()
        seg_rows = int(cls.estep_segment_rows)
        if seg_rows <= 0 or pack.n_barcodes == 0 or \
                not lib.dmx_estep_plan_supported(pack.n_genotypes, float(doublet_prior), cls._flavour()):
            return None
        cache = pack.__dict__.setdefault('_estep_plans', {})
        key = (seg_rows, bool(cls.schedule_barcodes))
        if key not in cache:
            dev = pack.device
            capacity = pack.n_barcodes + pack.n_rows // seg_rows + 1
            seg_prefix = torch.empty(pack.n_barcodes + 1, dtype=torch.int32, device=dev)
            item_slot = torch.empty(capacity, dtype=torch.int32, device=dev)
            ws_bytes = lib.dmx_estep_plan_workspace_bytes(pack.n_barcodes)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            h_items = C.c_int64(0)
            with torch.cuda.device(dev):
                _native.check(lib.dmx_estep_plan(
                    pack.barcode_offsets.data_ptr(), pack.barcode_order.data_ptr() if cls.schedule_barcodes else 0,
                    pack.n_barcodes, seg_rows, seg_prefix.data_ptr(), item_slot.data_ptr(), capacity, ws.data_ptr(),
                    ws_bytes, C.byref(h_items), _stream()), 'dmx_estep_plan')
            cache[key] = (seg_prefix, item_slot, int(h_items.value), seg_rows)
        return cache[key]

    @classmethod
    def _e_step(cls, pack: DevicePack, table: torch.Tensor, doublet_prior: float,
                prior_logits: Optional[torch.Tensor] = None, want_logits: bool = True, want_post: bool = True,
                want_singlets: bool = False, buffers: Optional[dict] = None):
        """(c) of north_star / demux.py:246-265 + softmax.  Returns (logits, posteriors, singlet_posteriors)."""
        lib = _native.load()
        dev = pack.device
        n_cols = n_options(pack.n_genotypes, doublet_prior)
        buffers = {} if buffers is None else buffers

        def buf(name, shape):
            t = buffers.get(name)
            if t is None or tuple(t.shape) != tuple(shape):
                t = torch.empty(shape, dtype=torch.float32, device=dev)
                buffers[name] = t
            return t

        logits = buf('logits', (pack.n_barcodes, n_cols)) if want_logits else None
        post = buf('post', (pack.n_barcodes, n_cols)) if want_post else None
        singlets = None
        if want_singlets:  # leading dimension padded to 4 (128-bit loads in the M-step); padding stays zero
            shape = (pack.n_barcodes, cls._table_ld(pack.n_genotypes))
            singlets = buffers.get('singlets')
            if singlets is None or tuple(singlets.shape) != shape:
                singlets = buffers['singlets'] = torch.zeros(shape, dtype=torch.float32, device=dev)
        plan = cls._estep_plan(pack, doublet_prior)
        seg_prefix, item_slot, n_items, seg_rows = plan if plan is not None else (None, None, 0, 0)
        workspace = None
        ws_bytes = lib.dmx_estep_workspace_bytes(pack.n_barcodes, pack.n_genotypes, float(doublet_prior), n_items,
                                                 1 if logits is None else 0)
        if ws_bytes > 0:
            workspace = buffers.get('estep_ws')
            if workspace is None or workspace.numel() < ws_bytes:
                workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                buffers['estep_ws'] = workspace
        with torch.cuda.device(dev):
            _native.check(lib.dmx_estep(
                pack.barcode_offsets.data_ptr(), pack.barcode_order.data_ptr() if cls.schedule_barcodes else 0,
                pack.csr_variant.data_ptr(), pack.csr_e.data_ptr(), pack.n_barcodes, table.data_ptr(), table.shape[1], pack.n_genotypes, float(doublet_prior),
                _native.ptr(prior_logits), n_cols,
                _native.ptr(logits), n_cols, _native.ptr(post), n_cols, _native.ptr(singlets),
                cls._table_ld(pack.n_genotypes),
                _native.ptr(workspace), ws_bytes, cls._flavour(), float(getattr(table, 'dmx_floor', 0.0)),
                _native.ptr(seg_prefix), _native.ptr(item_slot), n_items, seg_rows,
                _stream()), 'dmx_estep')
        return logits, post, singlets

    @classmethod
    def _mstep_plan(cls, pack: DevicePack):
        """Tier lists of the planned M-step (dmx_mstep_plan), computed once per pack; None when disabled."""
        if not cls.planned_mstep or pack.n_variants == 0:
            return None
        cached = pack.__dict__.get('_mstep_plan')
        if cached is None:
            lib = _native.load()
            dev = pack.device
            plan_bytes = lib.dmx_mstep_plan_bytes(pack.n_rows)
            plan = torch.empty(max(plan_bytes, 4), dtype=torch.uint8, device=dev)
            counts = (C.c_int64 * 3)()
            with torch.cuda.device(dev):
                _native.check(lib.dmx_mstep_plan(pack.variant_offsets.data_ptr(), pack.n_variants, pack.n_rows,
                                                 plan.data_ptr(), plan_bytes, counts, _stream()), 'dmx_mstep_plan')
            n_medium, n_heavy_variants, n_heavy_items = (int(c) for c in counts)
            scratch = torch.empty(max(n_heavy_items * pack.n_genotypes, 1), dtype=torch.float64, device=dev)
            cached = pack.__dict__['_mstep_plan'] = (plan, n_medium, n_heavy_variants, n_heavy_items, scratch)
        return cached

    @classmethod
    def _allreduce_tiles(cls, pack: DevicePack) -> int:
        tiles = int(cls.mstep_allreduce_tiles)
        if tiles <= 0:
            tiles = 1 if pack.n_variants * pack.n_genotypes * 4 < (1 << 30) else 4
        return max(1, min(tiles, max(pack.n_variants, 1)))

    @classmethod
    def _mstep_buffers(cls, pack: DevicePack) -> dict:
        """Output buffers of the (sharded) M-step: two float32 [v_pad, G] tables that alternate as `genotype_addition`
        (v_pad = V rounded up so that the table splits into equal per-rank slices, padding rows zero), plus, by exchange
        mode, the peer-mapped partial table or the float64 partials of the wide NCCL wire format.  Collective when
        sharded (the peer tables are mapped by all ranks together).  The CONTENTS of the two tables are unspecified
        (cached peer tables come back as the previous user left them): callers that read before writing zero them."""
        dev = pack.device
        world = 1
        assert cls.mstep_exchange in ('auto', 'peer', 'nccl'), f'unknown mstep_exchange {cls.mstep_exchange!r}'
        assert cls.mstep_allreduce_dtype in ('float32', 'float64'), f'unknown wire dtype {cls.mstep_allreduce_dtype!r}'
        if cls.process_group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(cls.process_group)
        G = pack.n_genotypes
        if world > 1 and cls.mstep_exchange in ('auto', 'peer'):
            import torch.distributed as dist
            from .distributed import peer_tables
            if dist.get_backend(cls.process_group) == 'nccl':
                v_pad = -(-max(pack.n_variants, 1) // (4 * world)) * (4 * world)
                peer = peer_tables(cls.process_group, dev, v_pad, G)
                if peer is not None:
                    return {'v_pad': v_pad, 'world': world, 'tables': peer['tables'], 'peer': peer}
            assert cls.mstep_exchange == 'auto', 'mstep_exchange = "peer" needs NCCL ranks with symmetric memory'
        v_pad = -(-max(pack.n_variants, 1) // world) * world
        out = {'v_pad': v_pad, 'world': world,
               'tables': [torch.zeros((v_pad, G), dtype=torch.float32, device=dev) for _ in range(2)]}
        if world > 1 and cls.mstep_allreduce_dtype == 'float64':
            n_tiles = cls._allreduce_tiles(pack)
            out['partial64'] = torch.zeros((v_pad, G), dtype=torch.float64, device=dev)
            out['slice64'] = torch.empty((-(-v_pad // n_tiles) + world) * G // world + G, dtype=torch.float64, device=dev)
        return out

    @classmethod
    def _m_step(cls, pack: DevicePack, singlets: torch.Tensor, out: Optional[torch.Tensor] = None,
                buffers: Optional[dict] = None) -> torch.Tensor:
        """(d) of north_star / demux.py:113-118 -> genotype_addition float32 [V, G] (+ cross-GPU sum when sharded).
        `out`: a [v_pad, G] table of _mstep_buffers (sharded) or any [V, G] float32 tensor (single GPU)."""
        lib = _native.load()
        dev = pack.device
        sharded = cls.process_group is not None
        V, G = pack.n_variants, pack.n_genotypes
        with torch.cuda.device(dev):
            plan = cls._mstep_plan(pack)
            blob, n_medium, n_heavy_variants, n_heavy_items, scratch = plan if plan is not None else (None, 0, 0, 0, None)
            if not sharded:
                if out is None:
                    out = torch.empty((V, G), dtype=torch.float32, device=dev)
                common = (pack.variant_offsets.data_ptr(), pack.csc_cb.data_ptr(), pack.csc_e.data_ptr(),
                          singlets.data_ptr(), singlets.shape[1], G, float(cls.contribution_power),
                          out.data_ptr(), G, 0, G, 0, V)
                if plan is None:
                    _native.check(lib.dmx_mstep(*common, _stream()), 'dmx_mstep')
                else:
                    _native.check(lib.dmx_mstep_planned(*common, blob.data_ptr(), pack.n_rows, n_medium,
                                                        n_heavy_variants, n_heavy_items, scratch.data_ptr(),
                                                        _stream()), 'dmx_mstep_planned')
                return out[:V]
            # (f) of north_star: the partial variant x genotype sums of all barcode shards are combined once per EM
            # iteration: dmx_mstep_allreduce computes the M-step over tiles of the variant range and NCCL sums tile k
            # (on the communicator's stream) while the kernel computes tile k + 1.
            from .distributed import native_comm
            if buffers is None:
                buffers = cls._mstep_buffers(pack)
            if out is None:
                out = buffers['tables'][0]
            peer = buffers.get('peer')
            if peer is not None:
                # own kernel over NVLink peer memory: local M-step into the peer-mapped partial, barrier, every rank
                # sums its slice over all partials and stores it into everybody's table, barrier
                partial, handle = peer['partial'], peer['handles'][0]
                common = (pack.variant_offsets.data_ptr(), pack.csc_cb.data_ptr(), pack.csc_e.data_ptr(),
                          singlets.data_ptr(), singlets.shape[1], G, float(cls.contribution_power),
                          partial.data_ptr(), G, 0, G, 0, V)
                if plan is None:
                    _native.check(lib.dmx_mstep(*common, _stream()), 'dmx_mstep')
                else:
                    _native.check(lib.dmx_mstep_planned(*common, blob.data_ptr(), pack.n_rows, n_medium,
                                                        n_heavy_variants, n_heavy_items, scratch.data_ptr(),
                                                        _stream()), 'dmx_mstep_planned')
                world = peer['world']
                assert out.data_ptr() in peer['pointers'], \
                    'the sharded M-step writes into one of the peer-mapped tables of _mstep_buffers (pass out=None or one of them)'
                ins = (C.c_void_p * world)(*peer['pointers'][partial.data_ptr()])
                outs = (C.c_void_p * world)(*peer['pointers'][out.data_ptr()])
                handle.barrier(channel=0, timeout_ms=60_000)
                _native.check(lib.dmx_peer_sum_f32(ins, outs, peer['rank'], world, partial.numel(), _stream()),
                              'dmx_peer_sum_f32')
                handle.barrier(channel=1, timeout_ms=60_000)
                buffers['last_exchange'] = cls.last_exchange = 'peer/float32'
                return out[:V]
            wide = cls.mstep_allreduce_dtype == 'float64'
            comm = native_comm(cls.process_group, dev)
            if comm is not None:
                _native.check(lib.dmx_mstep_allreduce(
                    pack.variant_offsets.data_ptr(), pack.csc_cb.data_ptr(), pack.csc_e.data_ptr(), singlets.data_ptr(),
                    singlets.shape[1], G, float(cls.contribution_power), out.data_ptr(),
                    _native.ptr(buffers.get('partial64')) if wide else 0,
                    _native.ptr(buffers.get('slice64')) if wide else 0, V, _native.ptr(blob), pack.n_rows, n_medium,
                    n_heavy_variants, n_heavy_items, _native.ptr(scratch), comm, cls._allreduce_tiles(pack),
                    1 if wide else 0, _stream()), 'dmx_mstep_allreduce')
                buffers['last_exchange'] = cls.last_exchange = \
                    f"nccl/{'float64' if wide else 'float32'}/tiles={cls._allreduce_tiles(pack)}"
                return out[:V]
            # process groups without NCCL (gloo over CUDA tensors): same algebra through torch.distributed
            import torch.distributed as dist
            partial = buffers['partial64'] if wide else out
            common = (pack.variant_offsets.data_ptr(), pack.csc_cb.data_ptr(), pack.csc_e.data_ptr(),
                      singlets.data_ptr(), singlets.shape[1], G, float(cls.contribution_power),
                      0 if wide else out.data_ptr(), G, partial.data_ptr() if wide else 0, G, 0, V)
            if plan is None:
                _native.check(lib.dmx_mstep(*common, _stream()), 'dmx_mstep')
            else:
                _native.check(lib.dmx_mstep_planned(*common, blob.data_ptr(), pack.n_rows, n_medium, n_heavy_variants,
                                                    n_heavy_items, scratch.data_ptr(), _stream()), 'dmx_mstep_planned')
            dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=cls.process_group)
            buffers['last_exchange'] = cls.last_exchange = f"torch.distributed/{'float64' if wide else 'float32'}"
            if wide:
                _native.check(lib.dmx_round_f64_to_f32(partial.data_ptr(), G, out.data_ptr(), G, V, G, _stream()),
                              'dmx_round_f64_to_f32')
        return out[:V]

    # ------------------------------------------------------------------------------------------------ aggregate_on_snps
    @classmethod
    def _snp_groups(cls, pack: DevicePack):
        """(barcode, SNP) groups of the matched molecule-level calls (demux.py:214-218), built once per pack:
        (grouped_variant, grouped_e, group_offsets, barcode_group_offsets, n_matched, n_groups)."""
        cached = pack.__dict__.get('_snp_groups')
        if cached is None:
            lib = _native.load()
            dev = pack.device
            n = pack.n_calls
            cap = max(n, 1)
            grouped_variant = torch.empty(cap, dtype=torch.int32, device=dev)
            grouped_e = torch.empty(cap, dtype=torch.float32, device=dev)
            group_offsets = torch.empty(cap + 1, dtype=torch.int64, device=dev)
            barcode_group_offsets = torch.empty(pack.n_barcodes + 1, dtype=torch.int64, device=dev)
            ws_bytes = lib.dmx_snp_groups_workspace_bytes(n)
            if ws_bytes < 0:
                _native.check(-1, 'dmx_snp_groups_workspace_bytes')
            workspace = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            h_matched, h_groups = C.c_int64(0), C.c_int64(0)
            lo, hi = 0, pack.n_barcodes  # barcode ids of a pack are local to its range
            with torch.cuda.device(dev):
                variant2snp = _to_device(np.ascontiguousarray(pack.variant2snp, dtype=np.int32), dev)
                _native.check(lib.dmx_build_snp_groups(
                    pack.call_variant.data_ptr(), pack.call_cb.data_ptr(), pack.call_e.data_ptr(), n,
                    variant2snp.data_ptr(), pack.n_snps, pack.n_barcodes, lo, hi, workspace.data_ptr(), ws_bytes,
                    grouped_variant.data_ptr(), grouped_e.data_ptr(), group_offsets.data_ptr(),
                    barcode_group_offsets.data_ptr(), C.byref(h_matched), C.byref(h_groups), _stream()),
                    'dmx_build_snp_groups')  # synchronises the stream: the pageable upload above is complete
            # FeatureLookup takes np.max of the features (utils.py:213): the reference raises on zero matched calls
            if h_matched.value == 0:
                raise ValueError('aggregate_on_snps=True needs at least one matched call')
            cached = pack.__dict__['_snp_groups'] = (grouped_variant, grouped_e, group_offsets, barcode_group_offsets,
                                                     int(h_matched.value), int(h_groups.value))
        return cached

    @classmethod
    def _e_step_aggregated(cls, pack: DevicePack, table: torch.Tensor, doublet_prior: float,
                           prior_logits: Optional[torch.Tensor] = None, want_post: bool = True,
                           want_singlets: bool = False, buffers: Optional[dict] = None):
        """`aggregate_on_snps = True` (demux.py:204-244) + softmax: (float64 logits, float64 posteriors, float32
        singlet posteriors for the M-step).  The reference feeds its float64 posteriors to the M-step; here they are
        rounded to float32 first (6e-8 relative on each weight)."""
        lib = _native.load()
        dev = pack.device
        n_cols = n_options(pack.n_genotypes, doublet_prior)
        buffers = {} if buffers is None else buffers
        grouped_variant, grouped_e, group_offsets, barcode_group_offsets, _n_matched, _n_groups = cls._snp_groups(pack)

        def buf(name, shape, dtype):
            t = buffers.get(name)
            if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
                t = buffers[name] = torch.zeros(shape, dtype=dtype, device=dev)
            return t

        logits = buf('snp_logits', (pack.n_barcodes, n_cols), torch.float64)
        post = buf('snp_post', (pack.n_barcodes, n_cols), torch.float64) if want_post else None
        singlets = buf('singlets', (pack.n_barcodes, cls._table_ld(pack.n_genotypes)), torch.float32) \
            if want_singlets else None
        ws_bytes = lib.dmx_snp_logits_workspace_bytes(pack.n_barcodes, n_cols)
        workspace = buffers.get('snp_ws')
        if workspace is None or workspace.numel() < ws_bytes:
            workspace = buffers['snp_ws'] = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _native.check(lib.dmx_snp_logits(
                barcode_group_offsets.data_ptr(), group_offsets.data_ptr(), grouped_variant.data_ptr(),
                grouped_e.data_ptr(), pack.n_barcodes, table.data_ptr(), table.shape[1], pack.n_genotypes,
                float(doublet_prior), float(cls.compensation_during_computing_barcode_logits), logits.data_ptr(),
                n_cols, workspace.data_ptr(), ws_bytes, _stream()), 'dmx_snp_logits')
            _native.check(lib.dmx_softmax_rows_f64(
                logits.data_ptr(), n_cols, _native.ptr(prior_logits), n_cols, pack.n_barcodes, n_cols,
                _native.ptr(post), n_cols, _native.ptr(singlets), cls._table_ld(pack.n_genotypes), pack.n_genotypes,
                _stream()), 'dmx_softmax_rows_f64')
        return logits, post, singlets

    @classmethod
    def _e_step_any(cls, pack, table, doublet_prior, **kwargs):
        """demux.py:193-202: the class flag picks the likelihood."""
        if cls.aggregate_on_snps:
            assert cls.process_group is None, 'aggregate_on_snps=True is a single-GPU path'
            kwargs.pop('want_logits', None)  # the float64 logits buffer is always filled
            return cls._e_step_aggregated(pack, table, doublet_prior, **kwargs)
        return cls._e_step(pack, table, doublet_prior, **kwargs)

    # ------------------------------------------------------------------------------------------------ public API
    @classmethod
    def predict_posteriors(cls, chromosome2compressed_snp_calls, genotypes, barcode_handler,
                           p_genotype_clip=0.01, doublet_prior=0.35):
        """One E-step with the given genotypes (demux.py:120-156): returns (logits_df, probs_df)."""
        pack = cls._pack_device(chromosome2compressed_snp_calls, genotypes, barcode_handler.n_barcodes,
                                add_data_prior=False, keep_calls=cls.aggregate_on_snps)
        table = cls._probs_table(pack, None, p_genotype_clip)
        table_is_finite = torch.isfinite(table).all()  # demux.py:135; read back together with the results
        logits, post, _ = cls._e_step_any(pack, table, doublet_prior)
        extras = [table_is_finite] + ([pack.betas_min] if pack.betas_min is not None else [])
        logits_np, post_np, finite, *betas_min = _to_host(logits, post, *extras)
        pack.check(*betas_min)  # demux.py:374 (the reference asserts before computing; the result is the same)
        assert bool(finite), 'non-finite genotype probabilities'
        names = pd.Index(option_names(genotypes.genotype_names, doublet_prior))
        index = pd.Index(list(barcode_handler.ordered_barcodes), name='BARCODE')
        # copy=False: the arrays were just created by the download and are owned by nothing else
        logits_df = pd.DataFrame(data=logits_np, index=index, columns=names, copy=False)
        probs_df = pd.DataFrame(data=post_np, index=index, columns=names, copy=False)
        return logits_df, probs_df

    @classmethod
    def _prior_logits_to_device(cls, barcode_prior_logits, n_barcodes: int, n_cols: int, dev):
        if barcode_prior_logits is None:
            return None
        assert barcode_prior_logits.shape == (n_barcodes, n_cols), 'wrong shape of priors'
        return _to_device(np.ascontiguousarray(barcode_prior_logits, dtype=np.float64), dev)

    @classmethod
    def staged_genotype_learning(cls, chromosome2compressed_snp_calls, genotypes, barcode_handler,
                                 n_iterations=5, p_genotype_clip=0.01, doublet_prior=0.,
                                 barcode_prior_logits: np.ndarray = None):
        """
        Generator over EM iterations (demux.py:68-118); yields (posterior DataFrame, debug dict with
        'barcode_logits', 'genotype_prior', 'genotype_addition') exactly like the reference, which means a
        device->host copy of the [B, C] matrices per iteration.  `learn_genotypes` avoids those copies.
        """
        assert 0 <= doublet_prior < 1
        assert not (cls.aggregate_on_snps and cls.process_group is not None), \
            'aggregate_on_snps=True is a single-GPU path'
        n_cols = n_options(genotypes.n_genotypes, doublet_prior)
        if barcode_prior_logits is not None:
            assert barcode_prior_logits.shape == (barcode_handler.n_barcodes, n_cols), 'wrong shape of priors'
        pack = cls._pack_device(chromosome2compressed_snp_calls, genotypes, barcode_handler.n_barcodes,
                                add_data_prior=True, keep_calls=cls.aggregate_on_snps)
        pack.check()
        prior_dev = cls._prior_logits_to_device(barcode_prior_logits, barcode_handler.n_barcodes, n_cols, pack.device)
        names = option_names(genotypes.genotype_names, doublet_prior)
        betas_host = pack.betas.cpu().numpy()
        betas_host.flags.writeable = False
        addition = torch.zeros_like(pack.betas)
        for iteration in range(n_iterations):
            table = cls._probs_table(pack, addition, p_genotype_clip)
            logits, post, singlets = cls._e_step_any(
                pack, table, doublet_prior, prior_logits=prior_dev if iteration == 0 else None, want_singlets=True)
            post_np, logits_np, addition_np = _to_host(post, logits, addition)
            post_df = pd.DataFrame(data=post_np, index=barcode_handler.ordered_barcodes, columns=names, copy=False)
            debug_information = {
                'barcode_logits': logits_np,
                'genotype_prior': betas_host,
                'genotype_addition': addition_np,
            }
            yield post_df, debug_information
            addition = cls._m_step(pack, singlets)

    @classmethod
    def learn_genotypes(cls, chromosome2compressed_snp_calls, genotypes, barcode_handler, n_iterations=5,
                        p_genotype_clip=0.01, doublet_prior=0., barcode_prior_logits: np.ndarray = None):
        """
        EM refinement of the genotypes (demux.py:34-66): returns (learnt genotypes, last posteriors).
        Same results as exhausting `staged_genotype_learning`, but everything stays on the device between
        iterations, only the singlet posteriors are materialised for the M-step, and the M-step after the last
        E-step (whose result the reference discards, demux.py:55,113) is not run.
        """
        assert 0 <= doublet_prior < 1
        assert not (cls.aggregate_on_snps and cls.process_group is not None), \
            'aggregate_on_snps=True is a single-GPU path'
        assert n_iterations >= 1, 'the reference unpacks the last generator item: n_iterations must be >= 1'
        n_cols = n_options(genotypes.n_genotypes, doublet_prior)
        if barcode_prior_logits is not None:
            assert barcode_prior_logits.shape == (barcode_handler.n_barcodes, n_cols), 'wrong shape of priors'
        pack = cls._pack_device(chromosome2compressed_snp_calls, genotypes, barcode_handler.n_barcodes,
                                add_data_prior=True, keep_calls=cls.aggregate_on_snps)
        pack.check()
        prior_dev = cls._prior_logits_to_device(barcode_prior_logits, barcode_handler.n_barcodes, n_cols, pack.device)
        post, addition = cls._em_iterations(pack, n_iterations, p_genotype_clip, doublet_prior, prior_dev)
        names = option_names(genotypes.genotype_names, doublet_prior)
        post_np, learnt_betas = _to_host(post, pack.raw_betas + addition)  # float32 add, demux.py:65
        post_df = pd.DataFrame(data=post_np, index=barcode_handler.ordered_barcodes, columns=names, copy=False)
        return genotypes._with_betas(learnt_betas), post_df

    @classmethod
    def _em_iterations(cls, pack: DevicePack, n_iterations: int, p_genotype_clip: float, doublet_prior: float,
                       prior_logits: Optional[torch.Tensor], want_post: bool = True):
        """Device-resident EM loop; returns (posteriors [B, C] of the last E-step, addition that fed it)."""
        buffers: dict = {}
        mbuf = cls._mstep_buffers(pack)
        V = pack.n_variants
        addition, spare = mbuf['tables']
        # genotype_addition starts at zero (demux.py:86).  The peer-mapped tables are cached per shape and REUSED by the
        # next EM run of the process, still holding that run's sums: without this reset the first E-step of a second
        # learn_genotypes call would see a stale addition (found by the 8-GPU lanes-vs-oracle check of bench.py).  Every
        # rank passed the closing barrier of the previous exchange before it gets here, so no peer writes are in flight.
        addition.zero_()
        table = None
        post = None
        for iteration in range(n_iterations):
            last = iteration == n_iterations - 1
            table = cls._probs_table(pack, addition[:V], p_genotype_clip, out=table)
            _, post, singlets = cls._e_step_any(
                pack, table, doublet_prior, prior_logits=prior_logits if iteration == 0 else None,
                want_logits=False, want_post=last and want_post, want_singlets=not last, buffers=buffers)
            if not last:
                cls._m_step(pack, singlets, out=spare, buffers=mbuf)
                spare, addition = addition, spare
        return post, addition[:V]


def _as_dtype(array: np.ndarray, dtype: np.dtype) -> np.ndarray:
    """Structured input with the expected packed layout passes through untouched; anything else is converted."""
    if array.dtype == dtype:
        return array
    out = np.empty(len(array), dtype=dtype)
    for name in dtype.names:
        out[name] = array[name]
    return out


def _foreign_hot_index(genotypes) -> dict:
    """hot_path_index() for a reference `ProbabilisticGenotypes` object (duck typing on var2varid)."""
    from .genotype_store import ProbabilisticGenotypes
    shim = ProbabilisticGenotypes.__new__(ProbabilisticGenotypes)
    shim.var2varid = genotypes.var2varid
    index = shim.hot_path_index()
    return index
