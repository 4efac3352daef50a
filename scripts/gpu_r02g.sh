#!/bin/bash
tag=${1:-r02g}
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q > $o/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $o/${tag}_pytest.log
tail -12 $o/${tag}_pytest.log | cut -c1-1800
timeout 600 python bench.py --configs c3 --no-cpu-baseline --e2e-steps 1 --steps 5 > $o/${tag}_c3.json 2> $o/${tag}_c3.err
python - <<PY
import json
d=json.load(open("$o/${tag}_c3.json"))
c=d["configs"]["c3"]
print("C2 back/fwd", d["roofline"]["kernel_ms"], d["roofline"]["forward"]["kernel_ms"], d["roofline"]["frac"], d["check"]["oracle"]["worst"])
print("C3", {k: c.get(k) for k in ("ms_per_iter","backward_ms","forward_ms","error")}, c.get("hbm_frac",{}).get("backward"), c.get("hbm_frac",{}).get("iteration"), c.get("oracle",{}).get("within_tolerance"))
PY
timeout 600 python scripts/perf_multi.py 65536 10 > $o/${tag}_multi10.json 2> $o/${tag}_multi.err; cat $o/${tag}_multi10.json; tail -3 $o/${tag}_multi.err
timeout 600 python scripts/perf_multi.py 65536 16 > $o/${tag}_multi16.json 2>> $o/${tag}_multi.err; cat $o/${tag}_multi16.json
timeout 600 python scripts/perf_multi.py 65536 6 > $o/${tag}_multi6.json 2>> $o/${tag}_multi.err; cat $o/${tag}_multi6.json
