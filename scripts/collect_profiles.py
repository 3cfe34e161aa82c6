"""
Turns the scratch artefacts in gpurun_out/ into the tracked evidence under profiles/ (run in the build container,
needs `ncu` only to read .ncu-rep files):

    python scripts/collect_profiles.py r01

  profiles/<round>_launches.csv           ncu launch list of `bench.py` (gpu__time_duration per launch)
  profiles/<round>_launch_shares.txt      the same aggregated per kernel (share of the step)
  profiles/<round>_ncu_<name>.txt         key metrics + stall reasons + opcode mix of an `ncu --set full` capture
  profiles/estep_traffic.json             dram bytes per launch of the pair E-step (read by bench.py -> roofline.traffic)
  profiles/<round>_*.log|json             microbenchmark, sweeps, parity report, e2e breakdown, bench lines
"""
import collections
import csv
import json
import re
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, SRC = ROOT / 'profiles', ROOT / 'gpurun_out'
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
OUT.mkdir(exist_ok=True)

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
]
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def ncu_csv(rep: Path, page: str):
    text = subprocess.run(['ncu', '-i', str(rep), '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(text.splitlines()))


def summarise_report(rep: Path, name: str):
    rows = ncu_csv(rep, 'raw')
    if len(rows) < 3:
        print('no kernels in', rep)
        return None
    header, units = rows[0], rows[1]
    lines, traffic = [], None
    for r in rows[2:]:
        d, u = dict(zip(header, r)), dict(zip(header, units))
        lines.append(f"== {d.get('Kernel Name', '?')}")
        for k in KEYS:
            if k in d:
                lines.append(f'{k:70s} {d[k]:>16s} {u.get(k, "")}')
        stalls = {k: float(v) for k, v in d.items()
                  if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio') and v}
        lines.append('stall reasons (warps per issue-active cycle): ' + ', '.join(
            f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}"
            for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))
        try:
            traffic = sum(float(d[k]) * UNIT[u[k]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
            lines.append(f'dram traffic per launch: {traffic / 1e6:.1f} MB')
        except Exception:  # noqa: BLE001
            pass
    src = ncu_csv(rep, 'source')
    hdr = [i for i, r in enumerate(src) if 'Source' in r and 'Instructions Executed' in r]
    if hdr:
        h = src[hdr[0]]
        si, ie, ws = h.index('Source'), h.index('Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
        count, stall = collections.Counter(), collections.Counter()
        for r in src[hdr[0] + 1:]:
            m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_\.]+)', r[si]) if len(r) > ie else None
            if m:
                try:
                    count[m.group(2).split('.')[0]] += int(r[ie])
                    stall[m.group(2).split('.')[0]] += int(r[ws])
                except ValueError:
                    pass
        total, stotal = sum(count.values()), max(sum(stall.values()), 1)
        lines.append(f'opcode mix of the first kernel ({total} warp instructions):')
        for op, n in count.most_common(14):
            lines.append(f'   {op:10s} {100 * n / total:5.1f} % of instructions   {100 * stall[op] / stotal:5.1f} % of stall samples')
    (OUT / f'{tag}_ncu_{name}.txt').write_text('\n'.join(lines) + '\n')
    return traffic


def launch_shares(path: Path):
    rows = [r for r in csv.reader(path.open()) if len(r) > 5]
    h = rows[0]
    ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1.0)
        tot[r[ki]][0] += 1
        tot[r[ki]][1] += v
    total = sum(v for _, v in tot.values())
    lines = [f'{"us total":>12s} {"share":>6s} {"launches":>8s}  kernel   (ncu serialises launches, cold caches: compare shares)']
    for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f'{v:12.1f} {100 * v / total:5.1f}% {n:8d}  {k[:110]}')
    (OUT / f'{tag}_launch_shares.txt').write_text('\n'.join(lines) + '\n')


if (SRC / 'launches.csv').exists():
    shutil.copy(SRC / 'launches.csv', OUT / f'{tag}_launches.csv')
    launch_shares(SRC / 'launches.csv')
for rep, name in (('prof_estep_full.ncu-rep', 'estep_pairs'), ('prof_aux.ncu-rep', 'mstep_singlets_table_softmax'),
                  ('prof_patch.ncu-rep', 'estep_pairs_patch_g200'), ('prof_singlets.ncu-rep', 'estep_singlets'),
                  ('r01b_mstep_light.ncu-rep', 'mstep_light_tier'), ('prof_mstep_tiers.ncu-rep', 'mstep_tiers'),
                  ('prof_small.ncu-rep', 'estep_lane_per_row_g4'), ('prof_snp_logits.ncu-rep', 'snp_logits')):
    if (SRC / rep).exists():
        traffic = summarise_report(SRC / rep, name)
        if name == 'estep_pairs' and traffic:
            (OUT / 'estep_traffic.json').write_text(json.dumps({
                'kernel': 'estep_pairs_warp_kernel', 'dram_bytes_per_launch': traffic, 'source': f'{tag}_ncu_{name}.txt',
                'workload': 'pbmc_32 scale 1.0 (R = 20.07 M rows, G = 32)'}) + '\n')
for f in ('microbench.log', 'microbench_packed_tile.log', 'bench_mstep.log', 'sweep_table.log', 'bench_snp_aggregate.log', 'parity_report_snp_aggregate.json', 'config4_shard.log', 'parity_report.json', 'profile_e2e.log', 'bench.log', 'bench_n2.log', 'bench_n8.log',
          'bench_reference.log', 'sweep_g32_c.log', 'sweep_g200_c.log', 'sanitizer.log', 'pytest_gpu.log',
          'pytest_dist.log'):
    if (SRC / f).exists():
        shutil.copy(SRC / f, OUT / f'{tag}_{f}')
print(sorted(p.name for p in OUT.iterdir()))
