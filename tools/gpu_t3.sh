#!/bin/bash
# usage: tools/gpu_t3.sh <tag>   parity suite + C3 (32- and 96-row chunks) + C4 bench lines
set -u
TAG=$1; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -12 $OUT/pytest.log
run() {
  name=$1; shift
  timeout 600 "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$name.json").read().strip().splitlines()[-1])
    print("$name", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("$name failed", e, open("$OUT/bench_$name.err").read()[-1500:])
PY
}
run c3 python bench.py --config c3 --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline
FM_STAGE_K=3 run c3_k3 python bench.py --config c3 --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline
run c4 python bench.py --config c4 --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline
