#!/bin/bash
# multi-GPU evidence on ONE box: bash scripts/gpu_multi.sh <tag> <N1> [N2 ...]   (the box must have max(N) GPUs)
tag=$1; shift
o=gpurun_out
mkdir -p $o
port=29500
for n in "$@"; do
  port=$((port+1))
  if [ "$n" = "1" ]; then
    timeout 300 python scripts/pcie_scaling.py > $o/${tag}_pcie_${n}gpu.json 2> $o/${tag}_pcie_${n}gpu.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port scripts/pcie_scaling.py > $o/${tag}_pcie_${n}gpu.json 2> $o/${tag}_pcie_${n}gpu.err
  fi
  tail -1 $o/${tag}_pcie_${n}gpu.json | cut -c1-700
done
for n in "$@"; do
  [ "$n" = "1" ] && continue
  port=$((port+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 > $o/${tag}_bench_${n}gpu.json 2> $o/${tag}_bench_${n}gpu.err
  python - <<PY
import json
try:
    d=json.loads(open("$o/${tag}_bench_${n}gpu.json").read().strip().splitlines()[-1])
    print("N=$n value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"), d["e2e"].get("ms_per_step"), "c5", {k: d["configs"].get("c5",{}).get(k) for k in ("ms_per_iter","c2_equivalent_iters_per_s","is_config_5","error","all_reduce")})
    print("   c5 oracle", d["configs"].get("c5",{}).get("oracle",{}).get("within_tolerance"), "e2e resident", d["e2e"].get("resident_inputs"))
except Exception as e:
    print("N=$n parse error", e); print(open("$o/${tag}_bench_${n}gpu.err").read()[-1500:])
PY
done
