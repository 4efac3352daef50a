#!/bin/bash
# e2e step time against the pipeline chunk (uniform chunks; 0 = the library's tapered default)
o=gpurun_out; mkdir -p $o
for c in 0 2368 0 2368; do
  timeout 300 python bench.py --configs "" --no-cpu-baseline --steps 3 --oracle-samples 0 --e2e-steps 5 --chunk $c > $o/e2e_chunk_$c.json 2> $o/e2e_chunk_$c.err
  python -c "
import json; d=json.load(open('$o/e2e_chunk_$c.json')); e=d['e2e']; print('chunk $c', 'e2e ms', round(e['ms_per_step'],2), 'it/s', round(e['value'],3), 'resident', round(e['resident_inputs']['ms_per_step'],2))"
done
