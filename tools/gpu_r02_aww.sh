#!/bin/bash
# Round 2: walls on the agent-warp mapping (N = 3, O = 3, W = 1 / 2): parity against the oracle, the golden fixtures and the
# group kernels bit for bit; bench at the C2 shape with walls, both mappings on the same box.
set -u
OUT=gpurun_out/${FM_OUT_TAG:-r02_aww}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "walls or mappings_bitwise or golden or w2 or w1" > $OUT/pytest_walls.log 2>&1; tail -5 $OUT/pytest_walls.log | cut -c1-400
b() { # tag, env-assignments, args
  tag=$1; shift
  timeout 400 python bench.py "$@" > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "%.4g" % d["value"], "us/step %.3f" % (1e3 * d["ms_per_step"]), "frac %.3f" % d["roofline"]["frac"], "closed %.3f" % (1e3 * d["closed_loop"]["ms_per_step"]), d["roofline"]["kernel"], "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print("$tag failed", e, open("$OUT/bench_$tag.err").read()[-1500:])
PY
}
b walls2_aw --walls 2 --steps 500 --warmup 25 --no-cpu-baseline --e2e-steps 3
b walls2_group --walls 2 --mapping group --steps 500 --warmup 25 --no-cpu-baseline --e2e-steps 3
b walls1_aw --walls 1 --steps 500 --warmup 25 --no-cpu-baseline --e2e-steps 3
b walls2_aw_driver --walls 2 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3
b c2_driver --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3
