#!/bin/bash
tag=${1:-r02x}
o=gpurun_out
mkdir -p $o
for ex in 0 1 2 3; do
  DDP_TILE_EXP=$ex timeout 600 python bench.py --configs "" --no-cpu-baseline --e2e-steps 1 --steps 6 --oracle-samples 8 > $o/${tag}_exp$ex.json 2> $o/${tag}_exp$ex.err
  python - <<PY
import json
d=json.load(open("$o/${tag}_exp$ex.json"))
print("EXP=$ex back", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "oracle worst", d["check"]["oracle"]["worst"], d["check"]["oracle"]["within_tolerance"], d["check"]["diverged"])
PY
done
