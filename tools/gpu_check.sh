#!/bin/bash
# Round-trip on the GPU box: parity tests, bench line, ncu launch list of one simulated day at peak prevalence.
# usage: tools/gpu_check.sh <tag> [full] [quick]   (full = also one `ncu --set full` capture of 6 launches;
#                                                    quick = skip the ensemble and drop-in tests)
tag=${1:-x}
sel="test"
if [ "$3" == "quick" ] || [ "$2" == "quick" ]; then sel="not ensemble and not dropin"; fi
timeout 900 python -m pytest tests -m gpu -k "${sel}" -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
tail -5 gpurun_out/pytest_$tag.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
timeout 300 python tools/profile_case.py --spinup 40 --days 2 > gpurun_out/plain_$tag.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
    -k regex:abm_ -s 720 -c 18 --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_case.py --spinup 40 --days 2 > gpurun_out/ncu_$tag.log 2>&1
if [ "$2" == "full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:abm_sweep -s 240 -c 6 -o gpurun_out/prof_$tag \
      python tools/profile_case.py --spinup 40 --days 2 > gpurun_out/ncu2_$tag.log 2>&1
fi
cat gpurun_out/plain_$tag.log
