"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line."""
import csv
import sys
from collections import defaultdict

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
h = rows[hdr_idx[0]]
end = hdr_idx[1] - 2 if len(hdr_idx) > 1 else len(rows)
col = {}
for i, n in enumerate(h):
    col.setdefault(n, i)
samples = defaultdict(float); insts = defaultdict(float); text = {}
stall_cols = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
stalls = defaultdict(lambda: defaultdict(float))
tot_s = tot_i = 0
for r in rows[hdr_idx[0] + 1:end]:
    if len(r) < len(h):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    text[line] = r[1]
    try:
        s = float(r[col['# Samples']] or 0); n = float(r[col['Instructions Executed']] or 0)
    except ValueError:
        continue
    samples[line] += s; insts[line] += n; tot_s += s; tot_i += n
    for c in stall_cols:
        v = r[col[c]]
        if v:
            stalls[line][c] += float(v)
print(f'total samples {tot_s:.0f} total warp instructions {tot_i:.0f}')
for line in sorted(samples, key=lambda k: -samples[k])[:top]:
    st = sorted(stalls[line].items(), key=lambda kv: -kv[1])[:3]
    print(f'{line:4d} {100 * samples[line] / tot_s:5.1f}% smp {100 * insts[line] / tot_i:5.1f}% inst  '
          f'{" ".join(f"{k[6:]}={v:.0f}" for k, v in st):40s} | {text[line].strip()[:90]}')
