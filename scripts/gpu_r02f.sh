#!/bin/bash
tag=${1:-r02f}
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q > $o/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $o/${tag}_pytest.log
tail -6 $o/${tag}_pytest.log | cut -c1-1500
for mb in 2 3; do
  DDP_SMALL_MINB=$mb timeout 600 python bench.py --configs c3 --no-cpu-baseline --e2e-steps 1 --steps 5 > $o/${tag}_c3_minb$mb.json 2> $o/${tag}_c3_minb$mb.err
  python - <<PY
import json
d=json.load(open("$o/${tag}_c3_minb$mb.json"))
c=d["configs"]["c3"]
print("C2 back/fwd", d["roofline"]["kernel_ms"], d["roofline"]["forward"]["kernel_ms"], d["roofline"]["frac"], d["check"]["oracle"]["worst"])
print("MINB=$mb", {k: c.get(k) for k in ("ms_per_iter","backward_ms","forward_ms","error")}, c.get("hbm_frac",{}).get("backward"), c.get("hbm_frac",{}).get("iteration"), c.get("oracle",{}).get("within_tolerance"))
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bp_small -s 1 -c 1 -f -o $o/${tag}_bp_small python bench.py --configs c3 --no-cpu-baseline --e2e-steps 1 --steps 3 --oracle-samples 0 > $o/${tag}_ncu1.log 2>&1
python scripts/ncu_summary.py $o/${tag}_bp_small.ncu-rep $o/${tag}_bp_small.txt > /dev/null 2>&1
head -34 $o/${tag}_bp_small.txt | tail -14
