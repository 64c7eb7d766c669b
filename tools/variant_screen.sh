#!/bin/bash
# One screening round trip: every library under variants/ timed on the same packed population at the epidemic peak and
# on a quiet day (tools/mode_times.py), then the bit-exact parity subset on one chosen variant.
# usage: tools/variant_screen.sh <tag> <parity-variant> "<peak list>" "<quiet list>"
tag=$1; pv=$2; peak=$3; quiet=$4
libs() { echo "$1" | tr ' ' '\n' | sed 's|^|variants/|; s|$|.so|' | paste -sd, -; }
timeout 300 python tools/mode_times.py --days 4 --libs "$(libs "$peak")" > gpurun_out/${tag}_variants_peak.txt 2>&1; tail -n +1 gpurun_out/${tag}_variants_peak.txt | cut -c1-400
timeout 300 python tools/mode_times.py --days 4 --spinup 5 --seed-fraction 0.00015 --libs "$(libs "$quiet")" > gpurun_out/${tag}_variants_quiet.txt 2>&1; cat gpurun_out/${tag}_variants_quiet.txt | cut -c1-400
ABM_B200_LIB=$PWD/variants/$pv.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_fixtures.py -x -q -k "not ensemble and not at_scale" > gpurun_out/${tag}_parity_$pv.log 2>&1; tail -3 gpurun_out/${tag}_parity_$pv.log
# optional 5th argument: a third list measured on a larger population (45 M agents: the event queues three times as long)
if [ -n "$5" ]; then
  timeout 300 python tools/mode_times.py --agents 45000000 --days 2 --libs "$(libs "$5")" > gpurun_out/${tag}_variants_45M.txt 2>&1; cat gpurun_out/${tag}_variants_45M.txt | cut -c1-400
fi
