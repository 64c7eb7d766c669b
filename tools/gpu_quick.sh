#!/bin/bash
# Quick round trip on the GPU box: parity tests (no ensembles), bench line (device-resident legs only), ncu launch list of
# one simulated day at high prevalence.   usage: tools/gpu_quick.sh <tag> [full]   (full = also `ncu --set full`)
tag=${1:-x}
timeout 900 python -m pytest tests -m gpu -x -q -k "not ensemble" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
tail -4 gpurun_out/pytest_$tag.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-ingest --full-run-days 0 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$tag.json'))
    r = d['roofline']
    print('ms/day', round(d['ms_per_step'], 4), 'sweep avg ms', round(r['avg_launch_ms'], 5), 'frac', round(r['frac'], 4), 'whole-day frac', round(r['whole_day_frac'], 4), {k: r.get(k) for k in ('event_kernel_avg_ms', 'table_kernel_avg_ms')})
except Exception as e:
    print('bench line unreadable', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
    -k regex:abm_ -s 720 -c 18 --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_case.py --spinup 40 --days 2 > gpurun_out/ncu_$tag.log 2>&1
python tools/launches.py gpurun_out/launches_$tag.csv | tail -20
if [ "$2" == "full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:abm_ -s 720 -c 9 -o gpurun_out/prof_$tag \
      python tools/profile_case.py --spinup 40 --days 2 > gpurun_out/ncu2_$tag.log 2>&1
fi
