#!/bin/bash
set -u
TAG=$1; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "assign or lexifair or parity" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
run() {
  name=$1; shift
  timeout 600 "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$name.json").read().strip().splitlines()[-1])
    print("$name", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"])
except Exception as e:
    print("$name failed", e, open("$OUT/bench_$name.err").read()[-1500:])
PY
}
for c in c3 c4; do
  run $c python bench.py --config $c --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline
  FM_LANES=1 run ${c}_l1 python bench.py --config $c --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline
  FM_LANES=4 run ${c}_l4 python bench.py --config $c --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline
done
python tools/bench_assign.py > $OUT/assign.jsonl 2>&1; tail -5 $OUT/assign.jsonl
