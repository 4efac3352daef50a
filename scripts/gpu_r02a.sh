#!/bin/bash
# first GPU pass of round 2: the whole GPU test-suite, the bench line (all configs), the reference arm
tag=${1:-r02a}
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $o/${tag}_smi.txt
timeout 1200 python -m pytest tests -m gpu -q -x > $o/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $o/${tag}_pytest.log
timeout 900 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err
echo "bench exit $?" >> $o/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err
tail -5 $o/${tag}_pytest.log
head -c 3000 $o/${tag}_bench.json
tail -5 $o/${tag}_bench.err
