#!/usr/bin/env python
"""
Benchmark of the demuxalot likelihood / EM hot path on B200 (contract: see the task statement / DESIGN.md section 7).

    python bench.py --gpus N --steps K --warmup W                    # ours, headline workload pbmc_32
    python bench.py --workload biobank_200 --gpus N ...              # north-star run as the headline (strong scaling)
    python bench.py --impl reference --gpus N --steps K ...          # CPU arm: the reference's algorithm (oracle port)

Headline (every N): BASELINE.json configs[1], per GPU -- synthetic PBMC-preprint scale, 32 donors (528 singlet +
doublet columns), 650k variants, 10k barcodes, ~20M read rows, `predict_posteriors` (doublet prior 0.35); weak scaling
(every rank its own 10k barcodes of the same donors; the E-step is barcode-local).  One "step" = one pass of the hot
path over the resident rows: probability table from betas + E-step over all R x C (row, column) pairs + row softmax.

Numbers on the JSON line:
  value        R*C*steps / device time of the steps (CUDA events per step, L2 flushed between steps, max over ranks)
  e2e          the same metric through the public API `Demultiplexer.predict_posteriors` with HOST inputs (warm), plus
               the cold first call of the process (index flattening, pageable copies, allocator growth)
  roofline     the dominant kernel (pair E-step) alone against the FP32 pipe that bounds it (and, for the record, the
               HBM formula of the contract)
  em           EM iterations/s of the learn_genotypes inner loop (table + E-step + M-step [+ cross-GPU sum])
  biobank_200  BASELINE.json configs[3], the north-star run: learn_genotypes, 200 donors / 100k barcodes / 5M variants /
               ~500M rows, 10 EM iterations, STRONG scaling over the N GPUs: sharded pack, EM loop including the
               cross-GPU sum, the exposed collective time, checksums that must not depend on N, and a barcode slice
               checked against the oracle in the same run
  em_32_3m     BASELINE.json configs[2] (N = 1): 10 EM iterations at 3M variants + learnt betas vs the oracle on a slice
  lanes_64     BASELINE.json configs[4] (N > 1): one 64-donor lane per GPU sharing the prior betas
  multi_gpu_parity (N > 1): sharded / lanes EM on a small data set against one GPU and the oracle
  cpu_baseline the oracle port (numpy, 1 core, as the reference is written) on a fixed barcode slice of the workload
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = 'read_row_x_column_updates_per_s'
UNIT = 'updates/s'
WORKLOAD = 'pbmc_32'
DOUBLET_PRIOR = 0.35
CPU_SAMPLE_BARCODES = 1024  # fixed sample of the CPU legs (first barcodes of the workload)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=WORKLOAD, choices=['pbmc_32', 'biobank_200', 'em_32_3m', 'lanes_64'])
    ap.add_argument('--scale', type=float, default=1.0, help='shrink barcodes/SNPs (debug only; invalid as a result)')
    ap.add_argument('--flavour', default='auto', choices=['auto', 'fast', 'exact'])
    ap.add_argument('--e2e-steps', type=int, default=None)
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='only the headline workload')
    ap.add_argument('--em-iterations', type=int, default=10)
    ap.add_argument('--parity-only', action='store_true', help='N > 1: only the multi_gpu_parity checks')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed regions run."""
    REASONS = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown', 0x2: 'applications_clocks_setting', 0x10: 'sync_boost'}

    def __init__(self, index: int, period_s: float = 0.02):
        self.samples, self.reasons, self.period, self.stop_flag = [], set(), period_s, threading.Event()
        self.max_mhz, self.thread, self.handle = None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # noqa: BLE001
            self.error = repr(exc)

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.handle is not None:
            self.stop_flag.clear()
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        return self

    def __exit__(self, *exc):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join()
            self.thread = None

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'note': 'no NVML samples'}
        return {'sm_mhz': statistics.median(self.samples), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def measured_peaks():
    path = ROOT / 'MEASURED_PEAKS.json'
    if path.exists():
        return float(json.loads(path.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def pin(array: np.ndarray) -> bool:
    """cudaHostRegister a numpy buffer so uploads are true async DMA from pinned memory."""
    import torch
    if array.nbytes == 0:
        return True
    rc = torch.cuda.cudart().cudaHostRegister(array.ctypes.data, array.nbytes, 0)
    return int(rc) == 0


def unpin(array: np.ndarray) -> None:
    """cudaHostUnregister: a registered buffer must be unregistered BEFORE numpy frees it -- otherwise the driver keeps
    the stale mapping and a later pageable copy out of recycled addresses fails with cudaErrorAlreadyMapped."""
    import torch
    if array.nbytes:
        torch.cuda.cudart().cudaHostUnregister(array.ctypes.data)


def slice_barcodes(ds, n_b: int):
    """First n_b barcodes of a dataset as a self-contained (calls, barcode_handler) pair."""
    from demuxalot_b200 import BarcodeHandler, CompressedSNPCalls
    calls = {}
    for chrom, c in ds.calls.items():
        mols = c.molecules[:c.n_molecules]
        sc = c.snp_calls[:c.n_snp_calls]
        keep_mol = mols['compressed_cb'] < n_b
        new_id = np.cumsum(keep_mol) - 1
        keep_call = keep_mol[sc['molecule_index']]
        sub = sc[keep_call].copy()
        sub['molecule_index'] = new_id[sub['molecule_index']]
        out = CompressedSNPCalls.__new__(CompressedSNPCalls)
        out.molecules, out.snp_calls = mols[keep_mol].copy(), sub
        out.n_molecules, out.n_snp_calls = len(out.molecules), len(sub)
        calls[chrom] = out
    return calls, BarcodeHandler(ds.barcode_handler.ordered_barcodes[:n_b])


def time_oracle_stages(calls, genotypes, handler, n_jobs: int, doublet_prior: float = DOUBLET_PRIOR):
    """One predict_posteriors pass of the oracle with the stages timed apart: (pack seconds, resident-step seconds
    [table + E-step + softmax on the packed rows], rows, columns)."""
    import oracle
    t0 = time.perf_counter()
    v2s, betas, _mol, rows = oracle.OracleDemultiplexer.pack_calls(calls, genotypes, False)
    t1 = time.perf_counter()
    table = oracle.probs_from_betas(v2s, betas, 0.01)
    logits = oracle.barcode_logits(rows['variant_id'], rows['compressed_cb'], rows['p_base_wrong'], table, doublet_prior,
                                   handler.n_barcodes, n_jobs=n_jobs)
    oracle.softmax_rows(logits)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, len(rows['variant_id']), logits.shape[1]


PARITY_FAILURES: list = []


def parity_check(ok: bool, what: str) -> None:
    """A failed parity check of an extra must not throw away the measured line -- and must not pass silently either:
    it is recorded, printed in the line (`parity_failures`) and turns the exit code non-zero after the line is out."""
    if not ok:
        PARITY_FAILURES.append(what)
        print(f'PARITY FAILURE: {what}', file=sys.stderr, flush=True)


def rel_err(got, want, floor=1e-3):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float((np.abs(got - want) / np.maximum(np.abs(want), floor)).max(initial=0))


# ------------------------------------------------------------------------------------------------- CPU arm
def run_reference(args, rank: int):
    """The reference's CPU algorithm (oracle port: the reference is Python and cannot travel to the GPU box) on a FIXED
    barcode sample of the headline workload, all host cores.  `value` is the resident step (rows already packed) so
    that it compares like with like with the GPU arm's `value`; `e2e` includes the pack."""
    if rank != 0:
        return
    from demuxalot_b200.synthetic import make_config
    import oracle
    cores = oracle.demux_oracle.default_n_jobs()
    n_b = CPU_SAMPLE_BARCODES if args.scale == 1.0 else 64
    ds = make_config('pbmc_32', scale=args.scale, n_barcodes=n_b)
    step_s, e2e_s = [], []
    rows = n_cols = 0
    for k in range(args.warmup + args.steps):
        t_pack, t_step, rows, n_cols = time_oracle_stages(ds.calls, ds.genotypes, ds.barcode_handler, cores)
        if k >= args.warmup:
            step_s.append(t_step)
            e2e_s.append(t_pack + t_step)
    ms = 1e3 * sum(step_s) / len(step_s)
    value = rows * n_cols / (ms / 1e3)
    e2e_value = rows * n_cols / (sum(e2e_s) / len(e2e_s))
    sample = (f'{n_b} barcodes ({rows} rows) of the pbmc_32 workload (same generator and seed), all {n_cols} columns; '
              f'value = table + E-step + softmax on packed rows, e2e = the same plus pack_calls; E-step columns sharded '
              f'over {cores} processes')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32 terms, f64 accumulate', 'data': 'synthetic',
        'config': {'workload': f'pbmc_32 (BASELINE.json configs[1]) -- fixed sample: {sample}',
                   'doublet_prior': DOUBLET_PRIOR},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0,
                'ms_per_step': 1e3 * sum(e2e_s) / len(e2e_s)},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- GPU arm
class Ctx:
    """Rank / device / collectives of this process."""

    def __init__(self, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.dev = torch.device('cuda', local_rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x: float, op) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def timed(self, fn, n_warm, n_steps, flush=None):
        """Per-step CUDA events on the launching stream; L2 flushed (untimed) between steps.  Milliseconds."""
        torch = self.torch
        for _ in range(n_warm):
            fn()
        self.barrier()
        events = []
        for _ in range(n_steps):
            if flush is not None:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            events.append((a, b))
        self.barrier()
        return [a.elapsed_time(b) for a, b in events]

    def wall(self, fn):
        """Wall time of fn() bracketed by barriers, max over ranks (seconds), and its result."""
        self.barrier()
        t0 = time.perf_counter()
        out = fn()
        self.torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return self.max(dt), out


def sm_info(lib, local_rank):
    info = [ctypes.c_int(0) for _ in range(3)]
    lib.dmx_device_info(local_rank, ctypes.byref(info[0]), ctypes.byref(info[1]), ctypes.byref(info[2]), None, None)
    return info[0].value


def em_loop_on_pack(ctx: Ctx, D, pack, n_iterations: int, doublet_prior: float):
    """Times the device-resident EM loop (table + E-step + M-step [+ cross-GPU sum]) with CUDA events; returns
    (seconds max over ranks, posteriors, addition)."""
    torch = ctx.torch
    ctx.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    post, addition = D._em_iterations(pack, n_iterations, 0.01, doublet_prior, None)
    b.record()
    ctx.barrier()
    return ctx.max(a.elapsed_time(b)) / 1e3, post, addition


def exchange_mode(D, pack) -> str:
    """How the M-step partials are summed across the ranks in this run (Demultiplexer.mstep_exchange resolved)."""
    if D.process_group is None:
        return 'single GPU'
    if 'peer' in D._mstep_buffers(pack):
        return 'peer: dmx_peer_sum_f32 (own reduce-scatter + all-gather kernel over NVLink peer memory, float32 partials)'
    return f'nccl: dmx_mstep_allreduce ({D.mstep_allreduce_dtype} wire, {D._allreduce_tiles(pack)} tile(s))'


def collective_exposed_ms(ctx: Ctx, D, pack, doublet_prior: float, reps: int = 3):
    """M-step alone vs M-step + cross-GPU sum on the same posteriors: (mstep_ms, mstep_allreduce_ms), max over ranks."""
    if ctx.world == 1:
        return None
    table = D._probs_table(pack, None, 0.01)
    _, _, singlets = D._e_step(pack, table, doublet_prior, want_logits=False, want_post=False, want_singlets=True)
    group, D.process_group = D.process_group, None
    alone = D._mstep_buffers(pack)
    local_ms = ctx.timed(lambda: D._m_step(pack, singlets, out=alone['tables'][0], buffers=alone), 1, reps)
    D.process_group = group
    del alone
    mbuf = D._mstep_buffers(pack)
    summed_ms = ctx.timed(lambda: D._m_step(pack, singlets, out=mbuf['tables'][0], buffers=mbuf), 1, reps)
    return ctx.max(statistics.mean(local_ms)), ctx.max(statistics.mean(summed_ms))


# ---- BASELINE configs[3]: the north-star run ---------------------------------------------------------------------------
def run_biobank(ctx: Ctx, D, args, scale: float):
    """learn_genotypes on the 200-donor biobank workload, strong scaling: the data set is the same for every N (device
    generator keyed by barcode), rank r starts with the calls of the barcodes b = r (mod N) (its share of the input),
    the sharded pack routes them to contiguous barcode ranges."""
    import oracle
    from demuxalot_b200 import BarcodeHandler
    from demuxalot_b200.demultiplexer import n_options
    from demuxalot_b200.synthetic_device import make_device_config
    torch, dist = ctx.torch, ctx.dist
    t0 = time.perf_counter()
    ds = make_device_config('biobank_200', scale=scale)
    t_host = time.perf_counter() - t0
    G, B = ds.genotypes.n_genotypes, ds.barcode_handler.n_barcodes
    C = n_options(G, DOUBLET_PRIOR)
    t0 = time.perf_counter()
    index = ds.genotypes.hot_path_index()
    t_index = time.perf_counter() - t0
    # the calls are resident in HBM; the genotype betas are the one host input of this pack (4 GB): pinned, like the
    # host inputs of the headline's e2e leg (pageable they upload at ~8 GB/s and dominate the pack at N = 8)
    betas_pinned = pin(ds.genotypes.variant_betas)
    mine = np.arange(ctx.rank, B, ctx.world)
    part = ds.device_calls(mine, ctx.dev)
    ctx.barrier()
    shard = (ctx.rank, ctx.world, dist.group.WORLD) if ctx.world > 1 else None
    if ctx.world > 1:
        D.process_group = dist.group.WORLD
    try:
        if ctx.world > 1:  # NCCL opens its point-to-point channels on first use: not part of a pack
            warm = torch.zeros(ctx.world * 4, dtype=torch.int32, device=ctx.dev)
            dist.all_to_all_single(torch.empty_like(warm), warm)
            dist.all_reduce(warm)
        pack_s, pack = ctx.wall(lambda: D._pack_device(None, ds.genotypes, B, add_data_prior=True, shard=shard,
                                                       device_parts=[part], keep_calls=False))
        # the same pack once more with a synchronisation after every stage (stage times; the sum is a little longer)
        pack_rows = pack.n_rows
        del pack
        torch.cuda.empty_cache()
        D.pack_profile = {}
        try:
            pack = D._pack_device(None, ds.genotypes, B, add_data_prior=True, shard=shard, device_parts=[part],
                                  keep_calls=False)
            pack_stages = {k: ctx.max(v) for k, v in sorted(D.pack_profile.items())}
        finally:
            D.pack_profile = None
        assert pack.n_rows == pack_rows
        del part
        torch.cuda.empty_cache()
        rows_total = ctx.sum(pack.n_rows)
        calls_total = ctx.sum(pack.n_calls)
        n_it = args.em_iterations
        D._em_iterations(pack, 1, 0.01, DOUBLET_PRIOR, None, want_post=False)  # plans, communicator, allocator
        em_s, post, addition = em_loop_on_pack(ctx, D, pack, n_it, DOUBLET_PRIOR)
        exposed = collective_exposed_ms(ctx, D, pack, DOUBLET_PRIOR)
        # N-independent fingerprints of the result (the data set does not depend on N; float64 regrouping of the
        # cross-GPU sum may move the last float32 bit of a few additions)
        learnt_sum = float(pack.raw_betas.sum(dtype=torch.float64).item() + addition.sum(dtype=torch.float64).item())
        post_singlet_mass = ctx.sum(float(post[:, :G].sum(dtype=torch.float64).item()))
        confident = ctx.sum(float((post.max(dim=1).values > 0.9).sum().item()))
        # M-step checksum property: column sums of the addition = sum over rows of (post (1 - e))^2
        lo, hi = pack.barcode_range
        out = {
            'workload': 'biobank_200 (BASELINE.json configs[3]): learn_genotypes, doublet prior 0.35, sharded by barcode',
            'scaling': 'strong', 'n_gpus': ctx.world, 'scale': scale,
            'donors': G, 'columns': C, 'variants': pack.n_variants, 'barcodes': B,
            'molecule_calls': int(calls_total), 'read_rows': int(rows_total), 'em_iterations': n_it,
            'host_genotypes_s': round(t_host, 2), 'host_index_s': round(t_index, 2),
            'pack_s': pack_s, 'pack_rows_per_s': rows_total / pack_s, 'pack_stages': pack_stages,
            'host_betas_pinned': bool(betas_pinned),
            'em_s': em_s, 'ms_per_iteration': 1e3 * em_s / n_it, 'iterations_per_s': n_it / em_s,
            'updates_per_s': rows_total * C * n_it / em_s,
            'what': 'pack = match + histogram + route + all-to-all + sort of the shard (inputs resident in HBM); '
                    'em = n_it x (table + E-step + M-step + cross-GPU sum), CUDA events, max over ranks',
            'learnt_betas_sum': learnt_sum, 'singlet_posterior_mass': post_singlet_mass,
            'barcodes_with_max_posterior_gt_0.9': int(confident),
            'peak_device_memory_gb': ctx.max(torch.cuda.max_memory_allocated() / 1e9),
            'mstep_exchange': exchange_mode(D, pack),
        }
        if exposed is not None:
            local_ms, summed_ms = exposed
            nbytes = pack.n_variants * G * 8 * (ctx.world - 1) / ctx.world  # float32: slice in from every peer, slice out to every peer
            out['collective'] = {'mstep_alone_ms': local_ms, 'mstep_plus_sum_ms': summed_ms,
                                 'exposed_ms': summed_ms - local_ms, 'bytes_on_wire_per_rank': int(nbytes),
                                 'exposed_frac_of_iteration': (summed_ms - local_ms) / (1e3 * em_s / n_it)}
        # size-independent property at full size: sum_v addition[v, g] == sum_rows (post_g (1 - e))^2 (this shard)
        table = D._probs_table(pack, None, 0.01)
        _, _, singlets = D._e_step(pack, table, DOUBLET_PRIOR, want_logits=False, want_post=False, want_singlets=True)
        saved, D.process_group = D.process_group, None
        part_add = D._m_step(pack, singlets)
        D.process_group = saved
        col, direct = 0, 0.0  # one column is enough for a checksum of checksums
        column = singlets[:, col].contiguous()
        for r0 in range(0, pack.n_rows, 1 << 26):
            r1 = min(pack.n_rows, r0 + (1 << 26))
            term = column[pack.csc_cb[r0:r1].long()] * (1.0 - pack.csc_e[r0:r1])  # float32, as demux.py:116-117
            direct += float((term * term).sum(dtype=torch.float64).item())
        got = float(part_add[:, col].sum(dtype=torch.float64).item())
        out['mstep_checksum_rel_err'] = abs(got - direct) / max(abs(direct), 1e-30)
        del table, singlets, part_add, column, term, post, addition, pack
    finally:
        D.process_group = None
    from demuxalot_b200.distributed import release_peer_tables
    release_peer_tables()
    torch.cuda.empty_cache()

    # barcode slice against the oracle, same run: public API on host-regenerated calls (rank 0)
    if ctx.rank == 0:
        n_slice = 16 if scale == 1.0 else 8
        ids = np.arange(n_slice)
        calls = ds.host_calls(ids)
        handler = BarcodeHandler(ds.barcode_handler.ordered_barcodes[:n_slice])
        t0 = time.perf_counter()
        gl, gp = D.predict_posteriors(calls, ds.genotypes, handler, doublet_prior=DOUBLET_PRIOR)
        t_gpu = time.perf_counter() - t0
        O = oracle.OracleDemultiplexer
        with all_host_cores():
            O.n_jobs = n_oracle_cores = oracle.demux_oracle.default_n_jobs()
            try:
                t0 = time.perf_counter()
                ol, op = O.predict_posteriors(calls, ds.genotypes, handler, doublet_prior=DOUBLET_PRIOR)
                t_cpu = time.perf_counter() - t0
            finally:
                O.n_jobs = 1
        # the device generator against its numpy mirror: identical records
        dev_part = ds.device_calls(ids, ctx.dev)
        host_rec = calls['chr1'].snp_calls[:calls['chr1'].n_snp_calls]
        same = bool(np.array_equal(dev_part['records'].cpu().numpy().reshape(-1), host_rec.view(np.uint8).reshape(-1)))
        out['slice_vs_oracle'] = {
            'barcodes': n_slice, 'rows_x_columns': int(gl.shape[1] * sum(c.n_snp_calls for c in calls.values())),
            'logits_rel_max': rel_err(gl.values, ol.values, 1e-30), 'posterior_abs_max': float(np.abs(gp.values - op.values).max()),
            'argmax_equal': bool((gp.values.argmax(1) == op.values.argmax(1)).all()),
            'device_generator_equals_host_mirror': same, 'gpu_s': round(t_gpu, 2), 'oracle_s': round(t_cpu, 2),
            'oracle_cores': n_oracle_cores}
        parity_check(out['slice_vs_oracle']['logits_rel_max'] <= 1e-5 and same, f"biobank_200 slice vs oracle: {out['slice_vs_oracle']}")
    ctx.barrier()
    if betas_pinned:
        unpin(ds.genotypes.variant_betas)
    return out


# ---- BASELINE configs[2] and configs[4] -------------------------------------------------------------------------------
def run_device_em(ctx: Ctx, D, args, name: str, scale: float, lanes: bool):
    """EM loop on a device-generated data set: `em_32_3m` (configs[2], one GPU) or one `lanes_64` lane per GPU
    (configs[4]: the lanes share donors and prior betas; M-step partials and molecule counts are summed across lanes)."""
    import oracle
    from demuxalot_b200 import BarcodeHandler
    from demuxalot_b200.demultiplexer import n_options
    from demuxalot_b200.synthetic_device import make_device_dataset
    torch, dist = ctx.torch, ctx.dist
    cfgs = {'em_32_3m': dict(n_genotypes=32, n_snps=1_500_000, n_barcodes=10_000, rows_per_barcode=3000,
                             unknown_genotype_fraction=0.78),
            'lanes_64': dict(n_genotypes=64, n_snps=650_000, n_barcodes=12_500, rows_per_barcode=2000)}
    cfg = dict(cfgs[name])
    if scale != 1.0:
        cfg['n_snps'] = max(64, int(cfg['n_snps'] * scale))
        cfg['n_barcodes'] = max(8, int(cfg['n_barcodes'] * scale))
    ds = make_device_dataset(seed=20260002 if name == 'em_32_3m' else 20260004,
                             calls_seed=ctx.rank if lanes else None, **cfg)
    G, B = ds.genotypes.n_genotypes, ds.barcode_handler.n_barcodes
    C = n_options(G, DOUBLET_PRIOR)
    part = ds.device_calls(np.arange(B), ctx.dev)
    if lanes and ctx.world > 1:
        D.process_group = dist.group.WORLD
    try:
        pack_s, pack = ctx.wall(lambda: D._pack_device(None, ds.genotypes, B, add_data_prior=True, device_parts=[part],
                                                       keep_calls=False))
        del part
        n_it = args.em_iterations
        D._em_iterations(pack, 1, 0.01, DOUBLET_PRIOR, None, want_post=False)
        em_s, post, addition = em_loop_on_pack(ctx, D, pack, n_it, DOUBLET_PRIOR)
        rows_total = ctx.sum(pack.n_rows)
        out = {'workload': {'em_32_3m': 'em_32_3m (BASELINE.json configs[2]): learn_genotypes, 32 donors, 3M variants',
                            'lanes_64': 'lanes_64 (BASELINE.json configs[4]): one 64-donor lane per GPU, shared prior '
                                        'betas, per-iteration cross-GPU sum'}[name],
               'n_gpus': ctx.world, 'scaling': 'weak', 'scale': scale, 'donors': G, 'columns': C,
               'variants': pack.n_variants, 'barcodes_per_gpu': B, 'read_rows': int(rows_total), 'em_iterations': n_it,
               'pack_s': pack_s, 'em_s': em_s, 'ms_per_iteration': 1e3 * em_s / n_it, 'iterations_per_s': n_it / em_s,
               'updates_per_s': rows_total * C * n_it / em_s,
               'learnt_betas_sum': float(pack.raw_betas.sum(dtype=torch.float64).item() + addition.sum(dtype=torch.float64).item())}
        exposed = collective_exposed_ms(ctx, D, pack, DOUBLET_PRIOR) if lanes else None
        if exposed is not None:
            out['collective'] = {'mstep_alone_ms': exposed[0], 'mstep_plus_sum_ms': exposed[1],
                                 'exposed_ms': exposed[1] - exposed[0]}
        del post, addition, pack
    finally:
        D.process_group = None
    from demuxalot_b200.distributed import release_peer_tables
    release_peer_tables()
    torch.cuda.empty_cache()
    if name == 'em_32_3m' and ctx.rank == 0:
        # learnt betas after the EM iterations against the oracle on a barcode slice (bar 1e-5 relative)
        n_slice = 48 if scale == 1.0 else 16
        calls = ds.host_calls(np.arange(n_slice))
        handler = BarcodeHandler(ds.barcode_handler.ordered_barcodes[:n_slice])
        learnt, gpost = D.learn_genotypes(calls, ds.genotypes, handler, n_iterations=n_it, doublet_prior=DOUBLET_PRIOR)
        O = oracle.OracleDemultiplexer
        with all_host_cores():
            O.n_jobs = oracle.demux_oracle.default_n_jobs()
            try:
                want, opost = O.learn_genotypes(calls, ds.genotypes, handler, n_iterations=n_it, doublet_prior=DOUBLET_PRIOR)
            finally:
                O.n_jobs = 1
        out['slice_vs_oracle'] = {'barcodes': n_slice, 'em_iterations': n_it,
                                  'learnt_betas_rel_max': rel_err(learnt.get_betas(), want.get_betas()),
                                  'posterior_abs_max': float(np.abs(gpost.values - opost.values).max())}
        parity_check(out['slice_vs_oracle']['learnt_betas_rel_max'] <= 1e-5, f"em_32_3m slice vs oracle: {out['slice_vs_oracle']}")
    ctx.barrier()
    return out


def run_multi_gpu_parity(ctx: Ctx, D):
    """Small data set: barcode-sharded EM and lane-per-GPU EM over the N ranks against one GPU and against the oracle
    (the driver's GPU test job has a single GPU, so the multi-GPU assertions live here as well)."""
    import oracle
    from demuxalot_b200 import BarcodeHandler, CompressedSNPCalls
    from demuxalot_b200.distributed import em_group, learn_genotypes_sharded
    from demuxalot_b200.synthetic import make_dataset
    O = oracle.OracleDemultiplexer
    ds = make_dataset(n_genotypes=12, n_snps=1500, n_barcodes=240, rows_per_barcode=150, seed=41)
    out = {}
    modes = [('peer', 'float32'), ('nccl', 'float64'), ('nccl', 'float32')]
    saved = (D.mstep_exchange, D.mstep_allreduce_dtype)
    for exchange, wire in modes:
        D.mstep_exchange, D.mstep_allreduce_dtype = ('auto' if exchange == 'peer' else 'nccl'), wire
        try:
            learnt, post = learn_genotypes_sharded(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=4,
                                                   doublet_prior=DOUBLET_PRIOR)
        finally:
            D.mstep_exchange, D.mstep_allreduce_dtype = saved
        ran = D.last_exchange or ''
        parity_check(ran.startswith(f'{exchange}/{wire}'), f'asked for the {exchange}/{wire} exchange, ran {ran!r}')
        wire = f'{exchange}_{wire}'
        if ctx.rank == 0:
            single, spost = D.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=4,
                                              doublet_prior=DOUBLET_PRIOR)
            want, opost = O.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=4,
                                            doublet_prior=DOUBLET_PRIOR)
            out[f'sharded_{wire}'] = {
                'exchange_that_ran': ran,
                'betas_rel_vs_single_gpu': rel_err(learnt.get_betas(), single.get_betas()),
                'betas_rel_vs_oracle': rel_err(learnt.get_betas(), want.get_betas()),
                'posterior_abs_vs_single_gpu': float(np.abs(post.values - spost.values).max()),
                'posterior_abs_vs_oracle': float(np.abs(post.values - opost.values).max())}
            parity_check(out[f'sharded_{wire}']['betas_rel_vs_oracle'] <= 1e-5, f'sharded EM ({wire}) vs oracle: {out[f"sharded_{wire}"]}')
    # the same call twice in one process: identical bits (the peer-mapped tables are cached and reused between EM runs;
    # a run must not see what the previous one left in them)
    first, first_post = learn_genotypes_sharded(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=4,
                                                doublet_prior=DOUBLET_PRIOR)
    again, again_post = learn_genotypes_sharded(ds.calls, ds.genotypes, ds.barcode_handler, n_iterations=4,
                                                doublet_prior=DOUBLET_PRIOR)
    same = bool(np.array_equal(np.array(first.get_betas()), np.array(again.get_betas())) and
                np.array_equal(first_post.values, again_post.values))
    out['repeat_call_bit_identical'] = same
    parity_check(same, 'a second learn_genotypes_sharded call in the same process gave different bits')
    # lanes: every rank its own barcodes; reference equivalent = one run over the union of the lanes
    lanes = [make_dataset(n_genotypes=12, n_snps=1500, n_barcodes=60, rows_per_barcode=150, seed=41, calls_seed=100 + r)
             for r in range(ctx.world)]
    mine = lanes[ctx.rank]
    with em_group():
        lane_learnt, lane_post = D.learn_genotypes(mine.calls, mine.genotypes, mine.barcode_handler, n_iterations=4,
                                                   doublet_prior=DOUBLET_PRIOR)
    if ctx.rank == 0:
        union = {}
        for chrom in lanes[0].calls:
            parts, offset = [], 0
            for lane in lanes:
                c = lane.calls[chrom]
                shifted = CompressedSNPCalls.__new__(CompressedSNPCalls)
                shifted.molecules = c.molecules[:c.n_molecules].copy()
                shifted.molecules['compressed_cb'] += offset
                shifted.snp_calls = c.snp_calls[:c.n_snp_calls].copy()
                shifted.n_molecules, shifted.n_snp_calls = c.n_molecules, c.n_snp_calls
                parts.append(shifted)
                offset += lane.barcode_handler.n_barcodes
            union[chrom] = CompressedSNPCalls.concatenate(parts)

        class UnionHandler:  # only n_barcodes / ordered_barcodes are read by the hot path (utils.py:60-66)
            n_barcodes = sum(lane.barcode_handler.n_barcodes for lane in lanes)
            ordered_barcodes = [f'{k}:{b}' for k, lane in enumerate(lanes) for b in lane.barcode_handler.ordered_barcodes]

        want, opost = O.learn_genotypes(union, lanes[0].genotypes, UnionHandler, n_iterations=4, doublet_prior=DOUBLET_PRIOR)
        n0 = lanes[0].barcode_handler.n_barcodes
        out['lanes'] = {'betas_rel_vs_oracle_on_union': rel_err(lane_learnt.get_betas(), want.get_betas()),
                        'posterior_abs_vs_oracle_lane0': float(np.abs(lane_post.values - opost.values[:n0]).max())}
        parity_check(out['lanes']['betas_rel_vs_oracle_on_union'] <= 1e-5, f"lanes EM vs oracle on the union: {out['lanes']}")
    ctx.barrier()
    return out


# ---- headline: BASELINE configs[1] ------------------------------------------------------------------------------------
def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device'
    torch.cuda.set_device(local_rank)
    ctx = Ctx(rank, local_rank, world)
    dev = ctx.dev
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from demuxalot_b200 import Demultiplexer, _native, build
    if rank == 0:
        build.build()
        from demuxalot_b200.synthetic_device import build_synth
        build_synth()
    if world > 1:
        dist.barrier()
    lib = _native.load()
    from demuxalot_b200.synthetic import make_config
    from demuxalot_b200.demultiplexer import n_options
    D = Demultiplexer
    D.estep_flavour = args.flavour
    if world > 1 and os.environ.get('DMX_HOST_CORES'):  # this rank's share of the host cores (partition_host_cores)
        D.host_gather_threads = max(1, min(16, int(os.environ['DMX_HOST_CORES']) // 2))
    sm_count = sm_info(lib, local_rank)

    if args.parity_only:
        assert world > 1, '--parity-only checks the multi-GPU paths'
        res = run_multi_gpu_parity(ctx, D)
        if rank == 0:
            print(json.dumps({'multi_gpu_parity': res, 'n_gpus': world, 'parity_failures': PARITY_FAILURES}))
        from demuxalot_b200.distributed import release_native_comms
        release_native_comms()
        dist.destroy_process_group()
        if PARITY_FAILURES:
            sys.exit(1)
        return

    if args.workload != 'pbmc_32':  # another workload as the headline
        sampler = ClockSampler(local_rank)
        with sampler:
            if args.workload == 'biobank_200':
                res = run_biobank(ctx, D, args, args.scale)
            else:
                res = run_device_em(ctx, D, args, args.workload, args.scale, lanes=args.workload == 'lanes_64')
        if rank == 0:
            line = {'metric': METRIC, 'value': res['updates_per_s'], 'unit': UNIT, 'n_gpus': world,
                    'steps': res['em_iterations'], 'warmup': 1, 'ms_per_step': res['ms_per_iteration'],
                    'higher_is_better': True, 'scaling': res['scaling'], 'vs_baseline': None,
                    'dtype': 'f32 terms, f64 accumulate', 'data': 'synthetic (device generator)',
                    'config': {'workload': res['workload'], 'step': 'one EM iteration: table + E-step + M-step + cross-GPU sum'},
                    'clocks': sampler.summary(), 'gpu_launches': 4 * res['em_iterations'], args.workload: res,
                    'parity_failures': PARITY_FAILURES}
            print(json.dumps(line))
        if world > 1:
            from demuxalot_b200.distributed import release_native_comms
            release_native_comms()
            dist.destroy_process_group()
        if PARITY_FAILURES:
            sys.exit(1)
        return

    # ---- cold first call: nothing cached, pageable host buffers ------------------------------------------------------
    ds = make_config('pbmc_32', scale=args.scale, calls_seed=rank)
    G = ds.genotypes.n_genotypes
    C = n_options(G, DOUBLET_PRIOR)
    B = ds.barcode_handler.n_barcodes
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    D.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=DOUBLET_PRIOR)
    torch.cuda.synchronize()
    cold_s = time.perf_counter() - t0
    ds.genotypes.__dict__.pop('_hot_index_cache', None)
    t0 = time.perf_counter()
    ds.genotypes.hot_path_index()  # flatten the var2varid dict once (host, cached; the reference redoes it per call)
    index_s = time.perf_counter() - t0
    pinned = all([pin(c.snp_calls) and pin(c.molecules) for c in ds.calls.values()] + [pin(ds.genotypes.variant_betas)])

    # ---- resident state ---------------------------------------------------------------------------------
    pack = D._pack_device(ds.calls, ds.genotypes, B, add_data_prior=False, keep_calls=False)
    R, V = pack.n_rows, pack.n_variants
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    buffers: dict = {}
    table = D._probs_table(pack, None, 0.01)

    def step():
        D._probs_table(pack, None, 0.01, out=table)
        D._e_step(pack, table, DOUBLET_PRIOR, want_logits=True, want_post=True, buffers=buffers)

    sampler = ClockSampler(local_rank)
    with sampler:
        step_ms = ctx.timed(step, args.warmup, args.steps, flush_buf)

        # the dominant kernel alone (no softmax outputs requested -> only the pair kernel is launched)
        def estep_only():
            D._e_step(pack, table, DOUBLET_PRIOR, want_logits=True, want_post=False, buffers=buffers)
        kernel_ms = ctx.timed(estep_only, args.warmup, args.steps, flush_buf)
        flavour_ms = {}
        for flavour in ('fast', 'exact'):  # both arithmetic flavours of the same step, for the record
            D.estep_flavour = flavour
            flavour_ms[flavour] = ctx.max(statistics.mean(ctx.timed(step, 2, max(3, args.steps // 4), flush_buf)))
        D.estep_flavour = args.flavour

        # EM inner loop: table(betas + addition) + E-step (singlet posteriors only) + M-step (+ cross-GPU sum)
        if world > 1:
            D.process_group = dist.group.WORLD
        mbuf = D._mstep_buffers(pack)
        em_exchange = exchange_mode(D, pack)
        em_state = {'k': 0}

        def em_iteration():
            cur, nxt = mbuf['tables'][em_state['k'] & 1], mbuf['tables'][(em_state['k'] + 1) & 1]
            D._probs_table(pack, cur[:V], 0.01, out=table)
            _, _, singlets = D._e_step(pack, table, DOUBLET_PRIOR, want_logits=False, want_post=False,
                                       want_singlets=True, buffers=buffers)
            D._m_step(pack, singlets, out=nxt, buffers=mbuf)
            em_state['k'] += 1
        em_ms = ctx.timed(em_iteration, args.warmup, args.steps, flush_buf)
        # the same iteration without the cross-GPU sum: what the exchange costs (weak-scaling efficiency of the EM loop)
        D.process_group = None
        del mbuf
        mbuf = D._mstep_buffers(pack)
        em_local_ms = ctx.timed(em_iteration, 2, max(5, args.steps // 2), flush_buf)
        del mbuf
        from demuxalot_b200.distributed import release_peer_tables
        release_peer_tables()

        # end to end through the public API, host inputs (pinned) -> host DataFrames
        # (the synthetic dataset holds millions of small Python objects -- var2varid keys, barcodes; a cyclic-GC pass
        # over them costs 20-200 ms and used to land inside a timed call at random: freeze them out of the collector)
        import gc
        gc.collect()
        gc.freeze()
        e2e_steps = args.e2e_steps or max(3, min(args.steps, 5))
        # warm-up calls keep their results alive like the timed loop does: two generations of pinned download
        # buffers are in use at any time, and both have to exist before the clock starts
        for _ in range(max(3, args.warmup)):
            logits_df, probs_df = D.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=DOUBLET_PRIOR)
        ctx.barrier()
        e2e_times = []
        for _ in range(e2e_steps):
            t0 = time.perf_counter()
            logits_df, probs_df = D.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler, doublet_prior=DOUBLET_PRIOR)
            torch.cuda.synchronize()
            e2e_times.append(time.perf_counter() - t0)
        ctx.barrier()
        del logits_df, probs_df

    total_ms = ctx.max(sum(step_ms))
    units_per_step = ctx.sum(float(R) * C)
    value = units_per_step * args.steps / (total_ms / 1e3)
    e2e_s = ctx.max(sum(e2e_times)) / e2e_steps
    e2e_value = units_per_step / e2e_s
    # bytes that cross PCIe per call: packed snp_calls records, the compressed_cb column of the molecules (gathered by
    # host threads into a pinned staging buffer; the full 12-byte records when host_gather_threads == 0) and the betas;
    # the sorted genotype keys / SNP index are derived from the genotypes object and cached on the device with it
    mol_bytes = 4 if D.host_gather_threads > 0 else 12
    h2d = sum(13 * c.n_snp_calls + mol_bytes * c.n_molecules for c in ds.calls.values()) + V * G * 4
    d2h = 2 * B * C * 4
    em_total_ms = ctx.max(sum(em_ms))

    kernel_s = ctx.max(statistics.mean(kernel_ms)) / 1e3
    hbm_peak, peak_kind = measured_peaks()
    algorithmic_bytes = R * (8 + 4 * G) + B * C * 4  # row records + gathered table rows + logits
    clocks = sampler.summary()
    sm_mhz = clocks['sm_mhz'] or 1965.0
    sm_max = clocks['sm_max_mhz'] or 1965.0
    upd_clk_sm = R * C / kernel_s / (sm_count * sm_mhz * 1e6)
    # FP32 pipe: 128 lane-ops/clk/SM (scalar FADD/FMUL at 4 warp-instructions/clk/SM, packed FADD2/FMUL2 at 2; measured
    # with scripts/microbench_packed_tile.cu, profiles/r02_microbench_packed_tile.log); one update = one add + one
    # multiply = 2 lane-ops -> 64 updates/clk/SM.  Expressed in TFLOP/s (an add or a multiply = one flop).
    pipe_peak_tflops = sm_count * 128 * sm_max * 1e6 / 1e12
    achieved_tflops = 2.0 * R * C / kernel_s / 1e12
    traffic, traffic_source = None, None
    traffic_file = ROOT / 'profiles' / 'estep_traffic.json'
    if traffic_file.exists():
        try:
            rec = json.loads(traffic_file.read_text())
            traffic, traffic_source = rec.get('dram_bytes_per_launch'), f"profiles/estep_traffic.json ({rec.get('source', 'ncu --set full')})"
        except Exception:  # noqa: BLE001
            pass

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32 terms, f64 accumulate', 'data': 'synthetic',
        'config': {
            'workload': 'pbmc_32 (BASELINE.json configs[1]): predict_posteriors, per GPU',
            'donors': G, 'columns': C, 'variants': V, 'barcodes_per_gpu': B, 'read_rows_per_gpu': R,
            'molecule_calls_per_gpu': pack.n_calls, 'doublet_prior': DOUBLET_PRIOR, 'estep_flavour': args.flavour,
            'step': 'probability table + E-step (all R x C pairs) + row softmax, rows resident in HBM',
            'l2': 'flushed between steps (256 MiB write, untimed); inputs 285 MB > L2 as well',
            'scale': args.scale, 'host_inputs_pinned': bool(pinned),
        },
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': 1e3 * e2e_s, 'steps': e2e_steps, 'ms_each': [round(1e3 * t, 2) for t in e2e_times],
                'api': 'Demultiplexer.predict_posteriors(host CompressedSNPCalls, genotypes, barcode_handler)',
                'host_gather_threads': int(D.host_gather_threads), 'gc': 'dataset objects frozen (gc.freeze)',
                'cold_first_call_ms': 1e3 * ctx.max(cold_s), 'cold_first_call_value': units_per_step / ctx.max(cold_s),
                'cold_note': 'first call of the process: pageable host buffers, genotype index not flattened yet '
                             f'(hot_path_index alone: {1e3 * index_s:.0f} ms), device index upload, allocator growth'},
        'gpu_launches': 3 * args.steps,
        'gpu_launches_note': 'per step: probs_table_vec4_kernel, estep_pairs_strip_kernel (row softmax fused for single-item '
                             'barcodes), softmax_rows_kernel (segment combine + softmax of the multi-item barcodes only)',
        'roofline': {
            'kernel': 'estep_pairs_strip_kernel', 'bound': 'fp32_pipe', 'achieved': achieved_tflops,
            'peak': pipe_peak_tflops, 'unit': 'TFLOP/s', 'frac': achieved_tflops / pipe_peak_tflops,
            'traffic': traffic, 'traffic_source': traffic_source, 'kernel_ms': kernel_s * 1e3,
            'peak_kind': f'{sm_count} SMs x 128 FP32 lane-ops/clk (measured issue rate of FADD2/FMUL2/FADD/FMUL, '
                         f'profiles/r02_microbench_packed_tile.log) x {sm_max:.0f} MHz; 2 flops per update',
            'updates_per_s': R * C / kernel_s, 'updates_per_clk_per_sm': upd_clk_sm,
            'fp32_pipe_ceiling_updates_per_clk_per_sm': 64.0,
            'hbm': {'algorithmic_bytes_per_launch': int(algorithmic_bytes), 'achieved_GBps': algorithmic_bytes / kernel_s / 1e9,
                    'peak_GBps': hbm_peak, 'frac': algorithmic_bytes / kernel_s / 1e9 / hbm_peak, 'peak_kind': peak_kind,
                    'note': '0.26 algorithmic bytes per update at G=32: the pair E-step is not HBM-bound'},
        },
        'flavours': {'step_ms': flavour_ms, 'value': {k: units_per_step / (v / 1e3) for k, v in flavour_ms.items()},
                     'note': 'same step with both E-step arithmetic flavours; |d posterior| distributions per config: '
                             'profiles/r02_parity_report.json'},
        'em': {'iterations_per_s': args.steps / (em_total_ms / 1e3), 'ms_per_iteration': em_total_ms / args.steps,
               'updates_per_s': units_per_step * args.steps / (em_total_ms / 1e3),
               'ms_per_iteration_without_exchange': ctx.max(statistics.mean(em_local_ms)),
               'what': 'table + E-step (singlet posteriors) + M-step' + (f' + cross-GPU sum ({em_exchange})' if world > 1 else '')},
    }
    del pack, table, buffers, flush_buf
    torch.cuda.empty_cache()

    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        n_b = CPU_SAMPLE_BARCODES if args.scale == 1.0 else 64
        calls, handler = slice_barcodes(ds, n_b)
        t_pack, t_step, rows_cpu, _ = time_oracle_stages(calls, ds.genotypes, handler, 1)
        line['cpu_baseline'] = {
            'value': rows_cpu * C / t_step, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'seconds': t_pack + t_step,
            'e2e_value': rows_cpu * C / (t_pack + t_step),
            'sample': f'first {n_b} barcodes ({rows_cpu} rows) of the same workload, all {C} columns, oracle (numpy, single '
                      f'process as the reference is written): value = table + E-step + softmax on packed rows '
                      f'({t_step:.1f} s), e2e_value adds pack_calls ({t_pack:.1f} s); host has {os.cpu_count()} logical cores',
        }
    if pinned:
        for c in ds.calls.values():
            unpin(c.snp_calls)
            unpin(c.molecules)
        unpin(ds.genotypes.variant_betas)
    del ds
    if not args.no_extras:
        extras_scale = args.scale
        sampler2 = ClockSampler(local_rank)
        with sampler2:
            biobank = run_biobank(ctx, D, args, extras_scale)
            line['biobank_200'] = biobank
            if world == 1:
                line['em_32_3m'] = run_device_em(ctx, D, args, 'em_32_3m', extras_scale, lanes=False)
            else:
                line['lanes_64'] = run_device_em(ctx, D, args, 'lanes_64', extras_scale, lanes=True)
                line['multi_gpu_parity'] = run_multi_gpu_parity(ctx, D)
        line['clocks_extras'] = sampler2.summary()
    line['parity_failures'] = PARITY_FAILURES
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        from demuxalot_b200.distributed import release_native_comms
        release_native_comms()
        dist.destroy_process_group()
    if PARITY_FAILURES:
        sys.exit(1)


ALL_HOST_CORES = None


class all_host_cores:
    """The CPU oracle legs run on rank 0 while the other ranks wait at a barrier: give them every host core back for
    that time (partition_host_cores otherwise leaves rank 0 with 1 / N of them)."""

    def __enter__(self):
        self.mine = None
        if ALL_HOST_CORES:
            try:
                self.mine = os.sched_getaffinity(0)
                os.sched_setaffinity(0, ALL_HOST_CORES)
            except (AttributeError, OSError):
                self.mine = None
        return self

    def __exit__(self, *exc):
        if self.mine:
            os.sched_setaffinity(0, self.mine)


def partition_host_cores(local_rank: int, local_world: int) -> int:
    """One contiguous block of the host cores per rank (ranks of a node otherwise pile their numpy / gather threads
    onto the same cores, and first-touch puts each rank's buffers on the memory node of the cores it runs on).
    Returns the number of cores of this rank."""
    global ALL_HOST_CORES
    try:
        cores = sorted(os.sched_getaffinity(0))
        ALL_HOST_CORES = set(cores)
        per = len(cores) // max(local_world, 1)
        if local_world > 1 and per >= 1:
            os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per])
            return per
        return len(cores)
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'ours' and world > 1:
        n_cores = partition_host_cores(local_rank, int(os.environ.get('LOCAL_WORLD_SIZE', world)))
        os.environ['DMX_HOST_CORES'] = str(n_cores)
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit(f'--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE is {world})')
    run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
