"""Executed warp instructions per opcode (and shared-memory wavefronts) of one kernel of an ncu report:
    python scripts/ncu_opcode_mix.py report.ncu-rep [kernel-id-filter e.g. :::1] [rows per unit, to normalise]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else ':::1'
unit = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hdr_i]
print(rows[0][1][:150] if rows and len(rows[0]) > 1 else '')
iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
iW, iWi = hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal')
cnt, samp, wf, wfi = (collections.Counter() for _ in range(4))
for r in rows[hdr_i + 1:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    m = re.match(r'(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS].strip())
    if not m:
        continue
    op = m.group(1).split('.')[0]
    if op in ('LDS', 'STS'):
        op = m.group(1)
    cnt[op] += int(r[iE]); samp[op] += int(r[iSamp] or 0); wf[op] += int(r[iW] or 0); wfi[op] += int(r[iWi] or 0)
tot = sum(cnt.values())
print(f'total warp instructions {tot}  ({tot / unit:.1f} per unit)   shared wavefronts {sum(wf.values())} ({sum(wf.values()) / unit:.1f} per unit, ideal {sum(wfi.values()) / unit:.1f})')
for op, c in cnt.most_common(28):
    print(f'{op:14s} {c / unit:9.1f}/unit {100 * c / tot:5.1f} %   stall samples {100 * samp[op] / max(sum(samp.values()), 1):5.1f} %   '
          f'wavefronts/unit {wf[op] / unit:7.1f} (ideal {wfi[op] / unit:7.1f})')
