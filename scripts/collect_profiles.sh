#!/bin/bash
# Round evidence, run on the GPU box:  bash scripts/collect_profiles.sh <tag>
# Writes everything under gpurun_out/<tag>_*; the summaries are then copied into profiles/ by scripts/summarise_profiles.py.
tag=${1:-r01b}
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $o/${tag}_smi.txt
timeout 900 python -m pytest tests -m gpu -q > $o/${tag}_pytest.log 2>&1
timeout 600 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err
timeout 300 python scripts/bench_configs.py c3 > $o/${tag}_c3.json 2>&1
timeout 300 python scripts/bench_configs.py c4 17760 > $o/${tag}_c4.json 2>&1
timeout 120 python __graft_entry__.py --smoke > $o/${tag}_smoke.log 2>&1
timeout 200 python scripts/perf_probe.py 9472 ltv > $o/${tag}_ltv.log 2>&1
timeout 300 python scripts/bench_configs.py solve c2 16384 > $o/${tag}_solve_c2.json 2>&1
timeout 300 python scripts/bench_configs.py solve c3 65536 > $o/${tag}_solve_c3.json 2>&1
./profiles/microbench/dmma_occ > $o/${tag}_dmma_occ.txt 2>&1
python scripts/pcie_probe.py > $o/${tag}_pcie.json 2>&1
# launch list of the bench command (kernel share of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bp_|fwd_|batch_stats|df_|kl_" -c 400 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $o/${tag}_launches_bench.log 2>&1
# full captures of the dominant kernels, one launch each
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bp_tile -c 1 -f -o $o/${tag}_bp_tile python scripts/perf_probe.py 65536 > $o/${tag}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_lin32x8_kernel -s 3 -c 1 -f -o $o/${tag}_fwd_lin \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $o/${tag}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kl_tile -c 1 -f -o $o/${tag}_kl_tile python scripts/bench_configs.py c4 9472 > $o/${tag}_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bp_small -s 1 -c 1 -f -o $o/${tag}_bp_small python scripts/bench_configs.py c3 262144 > $o/${tag}_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_pend_staged -s 2 -c 1 -f -o $o/${tag}_fwd_pend python scripts/bench_configs.py c3 262144 > $o/${tag}_ncu5.log 2>&1
# sanitizer on the new kernels (small shapes)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_misc.py tests/test_gpu_solve.py -q -x -k "kl or pend or multi or forward_costs" > $o/${tag}_memcheck.log 2>&1
echo "memcheck exit $?" >> $o/${tag}_memcheck.log
ls -la $o | tail -40
