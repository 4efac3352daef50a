#!/bin/bash
# Round evidence, run on the GPU box:  bash scripts/collect_profiles.sh <tag>
# Writes everything under gpurun_out/<tag>_*; scripts/summarise_profiles.py then copies the summaries into profiles/.
tag=${1:-r02}
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $o/${tag}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q > $o/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $o/${tag}_pytest.log
timeout 120 python __graft_entry__.py --smoke > $o/${tag}_smoke.log 2>&1
timeout 900 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err
timeout 600 python scripts/perf_lims.py 9472 0.3 > $o/${tag}_lims_tight.json 2> $o/${tag}_lims.err
timeout 600 python scripts/perf_lims.py 9472 3.0 > $o/${tag}_lims_loose.json 2>> $o/${tag}_lims.err
timeout 200 python scripts/perf_c1.py 16384 > $o/${tag}_c1.json 2>&1
timeout 200 python scripts/perf_single.py > $o/${tag}_single.json 2>&1
timeout 200 python scripts/perf_probe.py 9472 ltv > $o/${tag}_ltv.log 2>&1
timeout 300 python scripts/bench_configs.py solve c3 65536 > $o/${tag}_solve_c3.json 2>&1
python scripts/pcie_scaling.py > $o/${tag}_pcie_1gpu.json 2>&1
# launch list of the bench command (kernel share of the step; every configuration of the line)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bp_|fwd_|batch_stats|df_|kl_|peak_|commit_" -c 600 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --oracle-samples 0 > $o/${tag}_launches_bench.log 2>&1
# full captures of the dominant kernels, one launch each
cap() { # name regex skip cmd...
  n=$1; rx=$2; sk=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $sk -c 1 -f -o $o/${tag}_$n "$@" > $o/${tag}_ncu_$n.log 2>&1
  python scripts/ncu_summary.py $o/${tag}_$n.ncu-rep $o/${tag}_$n.txt > /dev/null 2>&1
}
cap bp_tile "bp_tile32x8_kernel" 3 python bench.py --configs "" --steps 2 --no-cpu-baseline --e2e-steps 1 --oracle-samples 0
cap fwd_lin "fwd_lin32x8_kernel" 4 python bench.py --configs "" --steps 2 --no-cpu-baseline --e2e-steps 1 --oracle-samples 0
cap bp_small "bp_small" 1 python bench.py --configs c3 --steps 3 --no-cpu-baseline --e2e-steps 1 --oracle-samples 0
cap fwd_pend "fwd_pend_staged" 2 python bench.py --configs c3 --steps 3 --no-cpu-baseline --e2e-steps 1 --oracle-samples 0
cap kl_tile "kl_tile" 1 python bench.py --configs c4 --steps 2 --no-cpu-baseline --e2e-steps 1 --oracle-samples 0 --batch 16384
cap kl_cached "kl_tile" 11 python bench.py --configs c4 --steps 2 --no-cpu-baseline --e2e-steps 1 --oracle-samples 0
cap bp_tile_lims "bp_tile32x8" 7 python scripts/perf_lims.py 2368 3.0
cap bp_tile_gps "bp_tile32x8_kernel<.*1, .0, .0, .0>|bp_tile32x8_kernelILb0ELb1" 0 python bench.py --configs c4 --steps 2 --no-cpu-baseline --e2e-steps 1 --oracle-samples 0 --batch 16384
# sanitizer on the kernels added or rewritten this round (small shapes)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_round2.py tests/test_gpu_back_pass.py -q -x -k "not boxqp_large or 24" > $o/${tag}_memcheck.log 2>&1
echo "memcheck exit $?" >> $o/${tag}_memcheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests/test_gpu_round2.py -q -x -k "tile32x8_boxqp or with_limits or regimes or covariance_cache or partial_state or warp_per_trajectory" > $o/${tag}_synccheck.log 2>&1
echo "synccheck exit $?" >> $o/${tag}_synccheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_round2.py tests/test_gpu_back_pass.py -q -x -k "small or tile32x8_boxqp or with_limits or regimes or covariance_cache or partial_state or warp_per_trajectory or boxqp_large and 24 or chunked" > $o/${tag}_racecheck.log 2>&1
echo "racecheck exit $?" >> $o/${tag}_racecheck.log
ls -la $o | grep ${tag}_ | tail -50
