#!/bin/bash
set -u
TAG=$1; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
run() {
  name=$1; shift
  timeout 600 "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$name.json").read().strip().splitlines()[-1])
    el = d.get("edge_list") or {}
    print("$name", d["value"], d["ms_per_step"], d["roofline"]["frac"], "edge", el.get("ms_per_step_with_edge_list"), el.get("ms_per_step_closed_loop"))
except Exception as e:
    print("$name failed", e, open("$OUT/bench_$name.err").read()[-1500:])
PY
}
run c3 python bench.py --config c3 --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline
FM_STAGE=k1x2 run c3_k1x2 python bench.py --config c3 --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline
run c3_again python bench.py --config c3 --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline
run c4 python bench.py --config c4 --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline
FM_STAGE=k3x1 run c4_k3x1 python bench.py --config c4 --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline
