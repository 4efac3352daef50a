#!/bin/bash
# ncu captures of (a) the three-CTAs-per-SM experiment of the backward tile kernel, (b) the limits branch with the warp-cooperative QP
tag=${1:-r02n}
o=gpurun_out
mkdir -p $o
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_tile32x8 -s 6 -c 1 -f -o $o/${tag}_bp_tile_exp4 python scripts/perf_tile_exp.py 65536 4 > $o/${tag}_ncu4.log 2>&1
python scripts/ncu_summary.py $o/${tag}_bp_tile_exp4.ncu-rep $o/${tag}_bp_tile_exp4.txt > /dev/null 2>&1
head -60 $o/${tag}_bp_tile_exp4.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_tile32x8 -s 7 -c 1 -f -o $o/${tag}_bp_tile_lims python scripts/perf_lims.py 2368 3.0 > $o/${tag}_ncul.log 2>&1
python scripts/ncu_summary.py $o/${tag}_bp_tile_lims.ncu-rep $o/${tag}_bp_tile_lims.txt > /dev/null 2>&1
head -40 $o/${tag}_bp_tile_lims.txt
