#!/usr/bin/env python
"""
Benchmark of the demuxalot likelihood / EM hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py --gpus N --steps K --warmup W            # ours
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference's algorithm (oracle port)

Workload at every N: BASELINE.json configs[1], per GPU -- synthetic PBMC-preprint scale, 32 donors (528 singlet +
doublet columns), 650k variants, 10k barcodes, ~20M read rows, `predict_posteriors` (doublet prior 0.35).
With N > 1 every rank holds its own 10k barcodes of the same donors (weak scaling; the E-step is barcode-local,
no data-path collective).  One "step" = one pass of the hot path over the resident rows: probability table from
betas + E-step over all R x C (row, column) pairs + row softmax.

Numbers on the JSON line:
  value        R*C*steps / device time of the steps (CUDA events per step, L2 flushed between steps, max over ranks)
  e2e          the same metric through the public API `Demultiplexer.predict_posteriors` with HOST inputs: upload of
               the packed count_snps records and betas, device row building, table, E-step, softmax, download of
               logits + posteriors, DataFrame assembly
  roofline     the dominant kernel (pair E-step) alone, against the measured HBM copy bandwidth
  em           EM iterations/s of the learn_genotypes inner loop (table + E-step + M-step [+ NCCL all-reduce])
  cpu_baseline the oracle port (numpy, 1 core, as the reference is written) on a barcode slice of the workload
"""
from __future__ import annotations

import argparse
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=WORKLOAD)
    ap.add_argument('--scale', type=float, default=1.0, help='shrink barcodes/SNPs (debug only; invalid as a result)')
    ap.add_argument('--flavour', default='fast', choices=['fast', 'exact'])
    ap.add_argument('--e2e-steps', type=int, default=None)
    ap.add_argument('--skip-cpu-baseline', action='store_true')
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


def time_oracle_slice(ds, n_b: int, n_jobs: int, repeats: int = 1):
    """Seconds per predict_posteriors call of the oracle on the first n_b barcodes, and the rows it covered."""
    import oracle
    calls, handler = slice_barcodes(ds, n_b)
    O = oracle.OracleDemultiplexer
    O.n_jobs = n_jobs
    try:
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            logits, _ = O.predict_posteriors(calls, ds.genotypes, handler, doublet_prior=DOUBLET_PRIOR)
            times.append(time.perf_counter() - t0)
        _, _, _, rows = O.pack_calls(calls, ds.genotypes, False)
    finally:
        O.n_jobs = 1
    return times, len(rows['variant_id']), logits.shape[1]


def calibrate_slice(ds, n_jobs: int, budget_s: float) -> int:
    """Number of leading barcodes whose oracle predict_posteriors call takes about budget_s (two-point fit)."""
    n1, n2 = 8, 48
    (t1,), _, _ = time_oracle_slice(ds, n1, n_jobs)
    (t2,), _, _ = time_oracle_slice(ds, n2, n_jobs)
    slope = max(t2 - t1, 1e-4) / (n2 - n1)
    fixed = max(t1 - n1 * slope, 0.0)
    n_b = int((budget_s - fixed) / slope) if budget_s > fixed else n2
    return max(n2, min(ds.barcode_handler.n_barcodes, n_b))


# ------------------------------------------------------------------------------------------------- CPU arm
def run_reference(args, rank: int):
    if rank != 0:
        return
    from demuxalot_b200.synthetic import make_config
    import oracle
    cores = oracle.demux_oracle.default_n_jobs()
    budget_s = 150.0  # whole run, so that the default invocation ends within a few minutes
    ds = make_config(args.workload, scale=args.scale, n_barcodes=4096 if args.scale == 1.0 else 64)
    n_b = calibrate_slice(ds, cores, budget_s / (args.steps + args.warmup))
    times, rows, n_cols = time_oracle_slice(ds, n_b, cores, repeats=args.warmup + args.steps)
    timed = times[args.warmup:]
    ms = 1e3 * sum(timed) / len(timed)
    value = rows * n_cols / (ms / 1e3)
    sample = (f'first {n_b} barcodes ({rows} rows) of the {args.workload} workload, all {n_cols} columns, '
              f'predict_posteriors end to end (pack + table + E-step + softmax), columns sharded over {cores} processes')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32 terms, f64 accumulate', 'data': 'synthetic',
        'config': {'workload': f'{args.workload} (BASELINE.json configs[1]) -- bounded sample', 'sample': sample,
                   'doublet_prior': DOUBLET_PRIOR},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from demuxalot_b200 import Demultiplexer, _native, build
    build.build()
    lib = _native.load()
    from demuxalot_b200.synthetic import make_config
    from demuxalot_b200.demultiplexer import n_options
    Demultiplexer.estep_flavour = args.flavour

    ds = make_config(args.workload, scale=args.scale, calls_seed=rank)
    G = ds.genotypes.n_genotypes
    C = n_options(G, DOUBLET_PRIOR)
    B = ds.barcode_handler.n_barcodes
    pinned = all([pin(c.snp_calls) and pin(c.molecules) for c in ds.calls.values()] + [pin(ds.genotypes.variant_betas)])
    ds.genotypes.hot_path_index()  # flatten the var2varid dict once (host, cached; the reference redoes it per call)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- resident state ---------------------------------------------------------------------------------
    pack = Demultiplexer._pack_device(ds.calls, ds.genotypes, B, add_data_prior=False)
    R, V = pack.n_rows, pack.n_variants
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    buffers: dict = {}
    table = Demultiplexer._probs_table(pack, None, 0.01)

    def step():
        Demultiplexer._probs_table(pack, None, 0.01, out=table)
        Demultiplexer._e_step(pack, table, DOUBLET_PRIOR, want_logits=True, want_post=True, buffers=buffers)

    def timed_loop(fn, n_warm, n_steps):
        """Per-step CUDA events on the launching stream; L2 flushed (untimed) between steps."""
        for _ in range(n_warm):
            fn()
        barrier()
        events = []
        for _ in range(n_steps):
            flush_buf.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            events.append((a, b))
        barrier()
        return [a.elapsed_time(b) for a, b in events]

    sampler = ClockSampler(local_rank)
    with sampler:
        step_ms = timed_loop(step, args.warmup, args.steps)

        # the dominant kernel alone (no softmax outputs requested -> only the pair kernel is launched)
        def estep_only():
            Demultiplexer._e_step(pack, table, DOUBLET_PRIOR, want_logits=True, want_post=False, buffers=buffers)
        kernel_ms = timed_loop(estep_only, args.warmup, args.steps)

        # EM inner loop: table(betas + addition) + E-step (singlet posteriors only) + M-step (+ all-reduce)
        if world > 1:
            Demultiplexer.process_group = dist.group.WORLD
        em_state = {'addition': torch.zeros_like(pack.betas), 'spare': torch.empty_like(pack.betas),
                    'spare64': torch.empty(pack.betas.shape, dtype=torch.float64, device=dev) if world > 1 else None}

        def em_iteration():
            Demultiplexer._probs_table(pack, em_state['addition'], 0.01, out=table)
            _, _, singlets = Demultiplexer._e_step(pack, table, DOUBLET_PRIOR, want_logits=False, want_post=False,
                                                   want_singlets=True, buffers=buffers)
            new = Demultiplexer._m_step(pack, singlets, out=em_state['spare'], out64=em_state['spare64'])
            em_state['spare'], em_state['addition'] = em_state['addition'], new
        em_ms = timed_loop(em_iteration, args.warmup, args.steps)
        Demultiplexer.process_group = None

        # end to end through the public API, host inputs (pinned) -> host DataFrames
        # (the synthetic dataset holds millions of small Python objects -- var2varid keys, barcodes; a cyclic-GC pass
        # over them costs 20-200 ms and used to land inside a timed call at random: freeze them out of the collector)
        import gc
        gc.collect()
        gc.freeze()
        e2e_steps = args.e2e_steps or max(3, min(args.steps, 5))
        # warm-up calls keep their results alive like the timed loop does: two generations of pinned download
        # buffers are in use at any time, and both have to exist before the clock starts (cudaHostAlloc of the
        # second 42 MB set used to land in the second timed call: +23 ms)
        for _ in range(max(3, args.warmup)):
            logits_df, probs_df = Demultiplexer.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler,
                                                                   doublet_prior=DOUBLET_PRIOR)
        barrier()
        e2e_times = []
        for _ in range(e2e_steps):
            t0 = time.perf_counter()
            logits_df, probs_df = Demultiplexer.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler,
                                                                   doublet_prior=DOUBLET_PRIOR)
            torch.cuda.synchronize()
            e2e_times.append(time.perf_counter() - t0)
        barrier()

    total_ms = reduce_max(sum(step_ms))
    units_per_step = reduce_sum(float(R) * C)
    value = units_per_step * args.steps / (total_ms / 1e3)
    e2e_s = reduce_max(sum(e2e_times)) / e2e_steps
    e2e_value = units_per_step / e2e_s
    # bytes that cross PCIe per call: packed snp_calls records, the compressed_cb column of the molecules (gathered by
    # host threads into a pinned staging buffer; the full 12-byte records when host_gather_threads == 0) and the betas;
    # the sorted genotype keys / SNP index are derived from the genotypes object and cached on the device with it
    mol_bytes = 4 if Demultiplexer.host_gather_threads > 0 else 12
    h2d = sum(13 * c.n_snp_calls + mol_bytes * c.n_molecules for c in ds.calls.values()) + V * G * 4
    d2h = 2 * B * C * 4
    em_total_ms = reduce_max(sum(em_ms))

    kernel_s = reduce_max(statistics.mean(kernel_ms)) / 1e3
    hbm_peak, peak_kind = measured_peaks()
    algorithmic_bytes = R * (8 + 4 * G) + B * C * 4  # row records + gathered table rows + logits
    traffic = None
    traffic_file = ROOT / 'profiles' / 'estep_traffic.json'
    if traffic_file.exists():
        try:
            traffic = json.loads(traffic_file.read_text()).get('dram_bytes_per_launch')
        except Exception:  # noqa: BLE001
            traffic = None
    achieved = algorithmic_bytes / kernel_s / 1e9
    import ctypes
    info = [ctypes.c_int(0) for _ in range(3)]
    lib.dmx_device_info(local_rank, ctypes.byref(info[0]), ctypes.byref(info[1]), ctypes.byref(info[2]), None, None)
    sm_count = info[0].value
    clocks = sampler.summary()
    sm_mhz = clocks['sm_mhz'] or 1965.0

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32 terms, f64 accumulate', 'data': 'synthetic',
        'config': {
            'workload': f'{args.workload} (BASELINE.json configs[1]): predict_posteriors, per GPU',
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
                'host_gather_threads': int(Demultiplexer.host_gather_threads), 'gc': 'dataset objects frozen (gc.freeze)'},
        'gpu_launches': 3 * args.steps,
        'gpu_launches_note': 'per step: probs_table_vec4_kernel, estep_pairs_warp_kernel, softmax_rows_kernel',
        'roofline': {
            'kernel': 'estep_pairs_warp_kernel', 'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
            'frac': achieved / hbm_peak, 'traffic': traffic, 'peak_kind': peak_kind,
            'algorithmic_bytes_per_launch': int(algorithmic_bytes), 'kernel_ms': kernel_s * 1e3,
            'note': 'the pair E-step is bound by the FP32 pipe, not by HBM (0.26 B per update at G=32); see alu',
            'alu': {
                'updates_per_s': R * C / kernel_s,
                'updates_per_clk_per_sm': R * C / kernel_s / (sm_count * sm_mhz * 1e6),
                # measured on this GPU type with scripts/microbench_packed_tile.cu (profiles/r01_microbench_packed_tile.log):
                # FADD2 / FMUL2 / FFMA2 issue at 2.0 warp-instructions/clk/SM in any mix (128 FP32 lane-ops/clk/SM, the
                # same lanes as scalar FP32); one update = one add + one multiply -> 64 updates/clk/SM
                'fp32_pipe_ceiling_updates_per_clk_per_sm': 64.0,
                'frac_of_fp32_pipe_ceiling': R * C / kernel_s / (sm_count * sm_mhz * 1e6) / 64.0,
                # 8 x 8 register tiles waste the lower halves of the diagonal tiles (528 of 640 slots useful at G=32)
                # and 30 of 32 lanes carry tiles: 64 * 528/640 * 30/32
                'tiling_ceiling_updates_per_clk_per_sm': 64.0 * (C / (64.0 * (((G + 7) // 8) * ((G + 7) // 8 + 1) // 2))) * (30 / 32 if (G + 7) // 8 == 4 else 1.0),
            },
        },
        'em': {'iterations_per_s': args.steps / (em_total_ms / 1e3), 'ms_per_iteration': em_total_ms / args.steps,
               'updates_per_s': units_per_step * args.steps / (em_total_ms / 1e3),
               'what': 'table + E-step (singlet posteriors) + M-step' + (' + NCCL all-reduce of f64 [V,G]' if world > 1 else '')},
    }

    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        n_b = calibrate_slice(ds, 1, 15.0)
        (t_cpu,), rows_cpu, _ = time_oracle_slice(ds, n_b, 1)
        line['cpu_baseline'] = {
            'value': rows_cpu * C / t_cpu, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'seconds': t_cpu,
            'sample': f'first {n_b} barcodes ({rows_cpu} rows) of the same workload, all {C} columns, oracle '
                      f'predict_posteriors end to end (numpy, single process as the reference is written); '
                      f'host has {os.cpu_count()} logical cores',
        }
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit(f'--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE is {world})')
    run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
