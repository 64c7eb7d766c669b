#!/bin/bash
# Round-end evidence on ONE B200: full GPU test suite, the bench line (both arms), the ncu launch list of one simulated day
# at high prevalence and one `ncu --set full` capture of the same launches, the per-day cost curve, the device timeline.
# usage: tools/final_measure.sh <tag>      (outputs under gpurun_out/, copied into profiles/ by tools/collect_profiles.py)
tag=${1:-r02}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err; tail -2 gpurun_out/${tag}_bench_1gpu.err
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
timeout 300 python tools/timeline.py --days 3 > gpurun_out/${tag}_timeline_1gpu.txt 2>&1
timeout 300 python tools/timeline.py --days 2 --spinup 5 --seed-fraction 0.00015 > gpurun_out/${tag}_timeline_1gpu_quiet_day.txt 2>&1
timeout 300 python tools/day_profile.py --days 365 --out gpurun_out/${tag}_day_profile.json > gpurun_out/${tag}_day_profile.txt 2>&1
timeout 300 python tools/profile_case.py --spinup 40 --days 2 > gpurun_out/${tag}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
    -k regex:abm_ -s 720 -c 18 --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_case.py --spinup 40 --days 2 > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.txt; tail -20 gpurun_out/${tag}_launches.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:abm_ -s 720 -c 18 -o gpurun_out/${tag}_ncu_full \
    python tools/profile_case.py --spinup 40 --days 2 > gpurun_out/${tag}_ncu_full.log 2>&1
python - <<PY
import json
d = json.loads(open('gpurun_out/${tag}_bench_1gpu.json').read().strip().splitlines()[-1])
r = d['roofline']
print('ms/day', round(d['ms_per_step'], 4), 'sweep avg ms', round(r['avg_launch_ms'], 5), 'frac', round(r['frac'], 4), 'whole-day', round(r['whole_day_frac'], 4),
      'e2e', d['e2e']['value'], 'ingest', d['config']['ingest_s'], 'worst day', d['config']['full_run']['worst_day'])
PY
