"""Aggregates the warp-state samples of an ncu report per CUDA source line.
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv ; python tools/ncu_hot_lines.py src.csv [N]
"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


per_line = defaultdict(lambda: [0, defaultdict(int), ''])
fname, hdr, total = '', None, 0
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        si = hdr.index('# Samples')
        stalls = [(i, c) for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0]:
        cur = (fname, int(r[0]))
        per_line[cur][2] = r[1]
    n = num(r[si])
    if n:
        per_line[cur][0] += n
        total += n
        for i, c in stalls:
            per_line[cur][1][c] += num(r[i])
print('total samples', total)
for key, (n, st, src) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    dom = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f'{key[0]}:{key[1]:5d} {n:7d} {100 * n / total:5.1f}%  {", ".join(f"{c[6:]}={v}" for c, v in dom):34s} {src.strip()[:100]}')
