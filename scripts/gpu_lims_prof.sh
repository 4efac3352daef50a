#!/bin/bash
# limits branch of the backward tile kernel: racecheck over its tests, then one ncu capture
tag=${1:-r02l}
o=gpurun_out
mkdir -p $o
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_round2.py tests/test_gpu_back_pass.py -q -x -k "tile32x8_boxqp or with_limits or inverted" > $o/${tag}_racecheck.log 2>&1
echo "racecheck exit $?" >> $o/${tag}_racecheck.log
tail -4 $o/${tag}_racecheck.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_tile32x8 -s 7 -c 1 -f -o $o/${tag}_bp_tile_lims python scripts/perf_lims.py 2368 3.0 > $o/${tag}_ncu.log 2>&1
python scripts/ncu_summary.py $o/${tag}_bp_tile_lims.ncu-rep $o/${tag}_bp_tile_lims.txt > /dev/null 2>&1
head -42 $o/${tag}_bp_tile_lims.txt | grep -E "time_duration|inst_executed.sum|stall_|dmma|registers"
