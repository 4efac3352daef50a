#!/bin/bash
# ncu capture of the cached-covariance KL kernel (kl_tile32x8_kernel<2>): 12th launch of a kl_tile kernel in bench.py --configs c4 --steps 2
tag=${1:-r02r}
o=gpurun_out
mkdir -p $o
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kl_tile -s 11 -c 1 -f -o $o/${tag}_kl_cached python bench.py --configs c4 --steps 2 --no-cpu-baseline --e2e-steps 1 --oracle-samples 0 > $o/${tag}_ncu_kl.log 2>&1
python scripts/ncu_summary.py $o/${tag}_kl_cached.ncu-rep $o/${tag}_kl_cached.txt > /dev/null 2>&1
head -60 $o/${tag}_kl_cached.txt
