#!/bin/bash
# usage: tools/gpu_t6.sh <tag>   rollout tests + config-5 rollout loop at a few batch sizes
set -u
TAG=$1; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "rollout or policy" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -25 $OUT/pytest.log
for B in 4096 65536 524288; do
  timeout 600 python bench.py --config c5 --envs $B --steps 50 > $OUT/bench_c5_$B.json 2> $OUT/bench_c5_$B.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_c5_$B.json").read().strip().splitlines()[-1])
    print("c5 B=$B", d["value"], d["ms_per_step"], "sim", d["simulator_share"], "mem", d["mem_gb"])
except Exception as e:
    print("c5 $B failed", e, open("$OUT/bench_c5_$B.err").read()[-2500:])
PY
done
