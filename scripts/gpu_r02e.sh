#!/bin/bash
tag=${1:-r02e}
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q > $o/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $o/${tag}_pytest.log
tail -8 $o/${tag}_pytest.log | cut -c1-1500
timeout 600 python bench.py --configs c3 --no-cpu-baseline --e2e-steps 1 --steps 5 > $o/${tag}_c3.json 2> $o/${tag}_c3.err
python - <<PY
import json
d=json.load(open("$o/${tag}_c3.json"))
c=d["configs"]["c3"]
print("C2 back/fwd", d["roofline"]["kernel_ms"], d["roofline"]["forward"]["kernel_ms"], d["roofline"]["frac"])
print("C3", {k: c.get(k) for k in ("ms_per_iter","backward_ms","forward_ms","error")}, c.get("hbm_frac",{}), c.get("oracle",{}).get("within_tolerance"))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bp_small -s 1 -c 1 -f -o $o/${tag}_bp_small python bench.py --configs c3 --no-cpu-baseline --e2e-steps 1 --steps 3 --oracle-samples 0 > $o/${tag}_ncu1.log 2>&1
python scripts/ncu_summary.py $o/${tag}_bp_small.ncu-rep $o/${tag}_bp_small.txt > /dev/null 2>&1
head -34 $o/${tag}_bp_small.txt | tail -28
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bp_tile -s 3 -c 1 -f -o $o/${tag}_bp_tile python bench.py --configs "" --no-cpu-baseline --e2e-steps 1 --steps 2 --oracle-samples 0 > $o/${tag}_ncu2.log 2>&1
python scripts/ncu_summary.py $o/${tag}_bp_tile.ncu-rep $o/${tag}_bp_tile.txt > /dev/null 2>&1
head -34 $o/${tag}_bp_tile.txt | tail -28
