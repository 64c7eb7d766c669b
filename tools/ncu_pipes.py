"""Pipe utilisation and warp-stall breakdown per launch from an `ncu --set full` report (no GPU needed).

    ncu -i report.ncu-rep --page raw --csv > raw.csv;  python tools/ncu_pipes.py raw.csv [kernel-name filter]

What bounds the sweep is not the issue rate as such but the ALU pipe (integer logic / compare / shift / select: half
rate, 16 lanes per clock and SM sub-partition): the table shows issue-slot, ALU, FMA(-heavy) and LSU utilisation side by
side with DRAM throughput, and the average number of warps per issued instruction waiting for each stall reason."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
flt = sys.argv[2] if len(sys.argv) > 2 else 'abm_'
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
PIPES = [('issue', 'sm__inst_issued.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct'),
         ('alu', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',),
         ('fma', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',),
         ('fmaheavy', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',),
         ('lsu', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',),
         ('dram', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed')]
STALLS = ['long_scoreboard', 'not_selected', 'math_pipe_throttle', 'wait', 'dispatch_stall', 'branch_resolving',
          'short_scoreboard', 'no_instruction', 'lg_throttle', 'mio_throttle', 'barrier']


def val(r, names):
    for n in names:
        if n in col and r[col[n]] not in ('', 'n/a'):
            try:
                return float(r[col[n]])
            except ValueError:
                pass
    return float('nan')


print('launch kernel                         us  | % of peak: ' + '  '.join(f'{p[0]:>8s}' for p in PIPES) +
      ' | warps per issue stalled on: ' + ' '.join(s[:9] for s in STALLS[:7]))
for i, r in enumerate(rows[2:]):
    name = r[col['Kernel Name']]
    if flt not in name:
        continue
    short = name.split('(')[0].replace('void ', '').replace('abm::', '')[:30]
    us = val(r, ['gpu__time_duration.sum'])
    pipes = '  '.join(f'{val(r, p[1:]):8.1f}' for p in PIPES)
    st = ' '.join(f'{val(r, [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"]):9.2f}' for s in STALLS[:7])
    print(f'{i:4d}   {short:30s} {us:6.1f} |            {pipes} |                              {st}')
