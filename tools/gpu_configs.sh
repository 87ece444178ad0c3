#!/bin/bash
# Diagnostic bench lines for the non-headline BASELINE configs.  usage: tools/gpu_configs.sh <tag>
set -u
TAG=$1; OUT=gpurun_out/$TAG; mkdir -p $OUT
for c in c1 c3 c4; do
  steps=300; [ $c = c1 ] && steps=5000
  timeout 600 python bench.py --config $c --steps $steps --warmup 25 --e2e-steps 5 > $OUT/bench_$c.json 2> $OUT/bench_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$c.json").read().strip().splitlines()[-1])
    print("$c", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
except Exception as e:
    print("$c failed", e, open("$OUT/bench_$c.err").read()[-1500:])
PY
done
