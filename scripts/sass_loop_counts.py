"""Opcode counts inside the main loop (the largest backward branch) of every kernel of a cuobjdump -sass listing.
   python scripts/sass_loop_counts.py <binary-or-.so> [name-filter]"""
import re
import subprocess
import sys

txt = subprocess.run(['cuobjdump', '-sass', sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ''
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n', 1)[0].strip()
    if flt not in name:
        continue
    lines = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', f)]
    best = None
    for addr, ins in lines:
        m = re.search(r'\bBRA\S*\s+(?:!?U?P\d+,\s+)?0x([0-9a-f]+)', ins)
        if m and int(m.group(1), 16) < addr:
            body = [i for a, i in lines if int(m.group(1), 16) <= a <= addr]
            if best is None or len(body) > len(best):
                best = body
    cnt = {}
    for ins in best or []:
        op = re.sub(r'^@!?U?P\d+\s+', '', ins).split()[0].split('.')[0]
        cnt[op] = cnt.get(op, 0) + 1
    print(name, 'loop instructions:', len(best or []), dict(sorted(cnt.items(), key=lambda kv: -kv[1])))
