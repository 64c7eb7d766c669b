"""Regenerate the committed ncu evidence from one `ncu --set full` capture of a simulated day's sub-step launches.

    ncu -i prof.ncu-rep --page details --csv > details.csv
    ncu -i prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py details.csv raw.csv <tag> ["title line"]

writes profiles/r01_ncu_full_details_<tag>.csv (copy), profiles/r01_ncu_summary_<tag>.txt (the scheduler /
occupancy / throughput lines per launch) and profiles/dram_traffic.json (dram bytes per launch, read by bench.py
for `roofline.traffic`).
"""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
details, raw, tag = sys.argv[1:4]
title = sys.argv[4] if len(sys.argv) > 4 else ''
WANT = ['Duration', 'DRAM Throughput', 'Compute (SM) Throughput', 'Issue Slots Busy', 'Executed Ipc Active',
        'Achieved Occupancy', 'Theoretical Occupancy', 'Registers Per Thread', 'Avg. Active Threads Per Warp',
        'No Eligible', 'L2 Hit Rate', 'L1/TEX Hit Rate', 'Warp Cycles Per Issued Instruction']

rows = list(csv.reader(open(details)))
h = rows[0]
c = {n: h.index(n) for n in ('ID', 'Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value')}
per = {}
for r in rows[1:]:
    if len(r) <= c['Metric Value'] or not r[c['ID']].isdigit():
        continue
    k = per.setdefault(int(r[c['ID']]), {'name': r[c['Kernel Name']], 'm': {}})
    k['m'].setdefault(r[c['Metric Name']], (r[c['Metric Value']], r[c['Metric Unit']]))
out = ['ncu --set full --clock-control none, one simulated day (6 sub-step launches) of 15 M agents at peak prevalence'
       + (('; ' + title) if title else ''),
       'kernel template ids: <259> night (INFECT|STEP|INCR), <135> go-out 08:00, <451> noon (dense, INCR), <75> go-home 16:00',
       '']
for i in sorted(per):
    out.append(f"[{i}] {per[i]['name'][:36]}")
    for n in WANT:
        if n in per[i]['m']:
            v, u = per[i]['m'][n]
            out.append(f'    {n:40s} {v} {u}')
    out.append('')
open(os.path.join(ROOT, 'profiles', f'r01_ncu_summary_{tag}.txt'), 'w').write('\n'.join(out))
shutil.copyfile(details, os.path.join(ROOT, 'profiles', f'r01_ncu_full_details_{tag}.csv'))

rr = list(csv.reader(open(raw)))
hh, units = rr[0], rr[1]


def col(name, want_unit):
    i = hh.index(name)
    scale = {'us': 1.0, 'ms': 1e3, 'ns': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'Kbyte': 1e-3, 'byte': 1e-6}[units[i]]
    return [float(r[i].replace(',', '')) * scale for r in rr[2:]]


us, rd, wr = col('gpu__time_duration.sum', 'us'), col('dram__bytes_read.sum', 'Mbyte'), col('dram__bytes_write.sum', 'Mbyte')
n = len(us)
json.dump({
    'bytes_per_launch': sum(a + b for a, b in zip(rd, wr)) * 1e6 / n,
    'kernel': 'abm_sweep_kernel',
    'launches': n,
    'avg_ns_under_ncu': sum(us) * 1e3 / n,
    'source': f'profiles/r01_ncu_full_details_{tag}.csv (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum, '
              f'mean over the {n} sub-step launches of one simulated day at 15 M agents: 3 night, go-out, noon, go-home)',
    'per_launch_mb': [{'us': a, 'read': b, 'write': w} for a, b, w in zip(us, rd, wr)],
}, open(os.path.join(ROOT, 'profiles', 'dram_traffic.json'), 'w'), indent=1)
print('\n'.join(out[:20]))
