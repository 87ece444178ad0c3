#!/bin/bash
# GPU pass for the group-kernel work: full parity suite, C3 / C4 / headline bench lines.  usage: tools/gpu_t1.sh <tag>
set -u
TAG=$1; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -12 $OUT/pytest.log
for c in c3 c4 c2; do
  steps=300; [ $c = c2 ] && steps=2000
  timeout 600 python bench.py --config $c --steps $steps --warmup 25 --e2e-steps 3 --no-cpu-baseline > $OUT/bench_$c.json 2> $OUT/bench_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$c.json").read().strip().splitlines()[-1])
    print("$c", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("$c failed", e, open("$OUT/bench_$c.err").read()[-1500:])
PY
done
