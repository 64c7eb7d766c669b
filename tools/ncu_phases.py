"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by phase of the sub-step kernel.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::1 > src.csv
    python tools/ncu_phases.py src.csv [path/to/abm_kernels.cuh]

Phases are found in the source itself: every `__device__` function and every `// ====... <title>` banner inside
abm_sweep_kernel starts one, so the ranges follow the code (the report must come from the same revision).
"""
import csv
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'covid19-agent-based-model_b200', 'csrc', 'abm_kernels.cuh')

marks = []          # (first line, name)
in_kernel = False
for no, line in enumerate(open(src), 1):
    m = re.match(r'\s*(?:template\s*<[^>]*>\s*)?__device__\s+(?:__forceinline__\s+)?[\w:<> ]+?\s+(\w+)\s*\(', line)
    if m:
        marks.append((no, 'fn ' + m.group(1)))
        in_kernel = False
    if re.match(r'abm_sweep_kernel\(', line):
        marks.append((no, 'kernel prologue'))
        in_kernel = True
    b = re.match(r'\s*// =+ (.+)$', line)
    if b and in_kernel:
        marks.append((no, b.group(1).strip()[:40]))
    if re.match(r'__global__', line) and not in_kernel:
        marks.append((no, 'other kernels'))
marks.sort()


def phase_of(line_no):
    name = 'other'
    for first, n in marks:
        if first <= line_no:
            name = n
        else:
            break
    return name


rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
h = rows[hdr[0]]
end = hdr[1] - 2 if len(hdr) > 1 else len(rows)
col = {}
for i, n in enumerate(h):
    col.setdefault(n, i)
ins, smp, thr = defaultdict(float), defaultdict(float), defaultdict(float)
ti = ts = 0.0
for r in rows[hdr[0] + 1:end]:
    if len(r) < len(h):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    try:
        n = float(r[col['Instructions Executed']] or 0)
        s = float(r[col['# Samples']] or 0)
        t = float(r[col['Thread Instructions Executed']] or 0) if 'Thread Instructions Executed' in col else 0.0
    except ValueError:            # a source line with commas inside a string literal (inline asm): skip it
        continue
    g = phase_of(line)
    ins[g] += n; smp[g] += s; thr[g] += t; ti += n; ts += s
for name in sorted(ins, key=lambda k: -ins[k]):
    if ins[name] or smp[name]:
        print(f'{name:42s} inst {100 * ins[name] / ti:5.1f}%  smp {100 * smp[name] / ts:5.1f}%  avg-active {thr[name] / max(ins[name], 1):5.1f}')
print('total warp instructions', ti)
