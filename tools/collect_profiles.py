"""Copy the evidence of tools/final_measure.sh from gpurun_out/ (scratch) into profiles/ (tracked), and derive the two
summaries the docs quote: the per-kernel ncu summary of one simulated day and profiles/dram_traffic.json.

    python tools/collect_profiles.py r02
"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
src, dst = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
for name in ('bench_1gpu.json', 'bench_reference_arm.json', 'timeline_1gpu.txt', 'timeline_1gpu_quiet_day.txt', 'day_profile.json',
             'day_profile.txt', 'launches.csv', 'launches.txt', 'pytest_gpu.log'):
    p = os.path.join(src, f'{tag}_{name}')
    if os.path.exists(p):
        shutil.copy(p, os.path.join(dst, f'{tag}_{name}'))

# ---- DRAM traffic of the sweep kernel per launch (launch list: one row per metric)
rows = [r for r in csv.reader(open(os.path.join(src, f'{tag}_launches.csv'))) if len(r) > 10 and r[0].isdigit()]
d = defaultdict(dict)
for r in rows:
    d[int(r[0])]['name'] = r[4]
    d[int(r[0])][r[12]] = float(r[14].replace(',', ''))
sweeps = [x for i, x in sorted(d.items()) if 'abm_sweep_kernel' in x['name']]
per = [dict(kernel=x['name'].split('(')[0].replace('void ', ''), us=x['gpu__time_duration.sum'] / 1e3, read=x['dram__bytes_read.sum'] / 1e6,
            write=x['dram__bytes_write.sum'] / 1e6, warp_inst_M=x['smsp__inst_executed.sum'] / 1e6) for x in sweeps[:6]]
traffic = dict(bytes_per_launch=sum((p['read'] + p['write']) * 1e6 for p in per) / len(per), kernel='abm_sweep_kernel', launches=len(per),
               avg_ns_under_ncu=sum(p['us'] for p in per) / len(per) * 1e3,
               source=f'profiles/{tag}_launches.csv', per_launch_mb=per)
json.dump(traffic, open(os.path.join(dst, 'dram_traffic.json'), 'w'), indent=1)
print('dram_traffic.json: mean', round(traffic['bytes_per_launch'] / 1e6, 1), 'MB per sweep launch')

# ---- ncu --set full summary per kernel
rep = os.path.join(src, f'{tag}_ncu_full.ncu-rep')
if os.path.exists(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'details', '--csv'], capture_output=True, text=True).stdout
    want = ['Duration', 'DRAM Throughput', 'Compute (SM) Throughput', 'Issue Slots Busy', 'Executed Ipc Active', 'Achieved Occupancy',
            'Theoretical Occupancy', 'Registers Per Thread', 'Avg. Active Threads Per Warp', 'No Eligible', 'L2 Hit Rate', 'L1/TEX Hit Rate',
            'Warp Cycles Per Issued Instruction', 'Memory Throughput']
    byid = {}
    for r in csv.DictReader(out.splitlines()):
        byid.setdefault((int(r['ID']), r['Kernel Name']), {})[r['Metric Name']] = (r['Metric Value'], r['Metric Unit'])
    lines = [f'ncu --set full --clock-control none, one simulated day (6 sub-steps x [sweep, event kernel, table kernel]) of 15 M agents at high',
             'prevalence (tools/profile_case.py --spinup 40).  Sweep template ids: <259> night (INFECT|STEP|INCR), <135> go-out 08:00,',
             '<451> noon (dense, INCR), <75> go-home 16:00.  Per-launch times under ncu are cold-cache and serialised.', '']
    for (i, k), m in sorted(byid.items()):
        lines.append(f'[{i}] {k[:60]}')
        for w in want:
            if w in m:
                lines.append(f'    {w:40s} {m[w][0]} {m[w][1]}')
        lines.append('')
    open(os.path.join(dst, f'{tag}_ncu_summary.txt'), 'w').write('\n'.join(lines))
    print('wrote', f'profiles/{tag}_ncu_summary.txt', len(byid), 'launches')
