#!/bin/bash
# Short round-end evidence on ONE B200 when GPU minutes are scarce: the GPU suite without its long members (ensembles,
# BASELINE-size parity), the bench line, the ncu launch list of one simulated day, the device timelines, and one
# `ncu --set full` capture of the same launches (last: it is the first thing to drop).   usage: tools/final_measure_short.sh <tag>
tag=${1:-r02f}
timeout 300 python -m pytest tests -m gpu -x -q -k "not ensemble and not at_scale" > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err; tail -2 gpurun_out/${tag}_bench_1gpu.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${tag}_bench_1gpu.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('ms/day', round(d['ms_per_step'], 4), 'sweep avg ms', round(r['avg_launch_ms'], 5), 'frac', round(r['frac'], 4), 'whole-day', round(r['whole_day_frac'], 4),
          'events', r.get('event_kernel_avg_ms'), 'table', r.get('table_kernel_avg_ms'), 'e2e', d['e2e']['value'], d['e2e']['seconds'], 'full run s', d['config']['full_run']['seconds'],
          'worst day', d['config']['full_run']['worst_day'], 'parity', d.get('parity_check'))
except Exception as e:
    print('bench line unreadable', e)
PY
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
    -k regex:abm_ -s 720 -c 18 --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_case.py --spinup 40 --days 2 > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.txt; tail -20 gpurun_out/${tag}_launches.txt
timeout 100 python tools/timeline.py --days 3 > gpurun_out/${tag}_timeline_1gpu.txt 2>&1; tail -1 gpurun_out/${tag}_timeline_1gpu.txt
timeout 100 python tools/timeline.py --days 2 --spinup 5 --seed-fraction 0.00015 > gpurun_out/${tag}_timeline_1gpu_quiet_day.txt 2>&1; tail -1 gpurun_out/${tag}_timeline_1gpu_quiet_day.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:abm_ -s 720 -c 18 -o gpurun_out/${tag}_ncu_full \
    python tools/profile_case.py --spinup 40 --days 2 > gpurun_out/${tag}_ncu_full.log 2>&1; tail -2 gpurun_out/${tag}_ncu_full.log
