#!/bin/bash
# Round 2, GPU session U: next-episode prefetch for the agent-warp mapping (and under stream capture).
set -u
OUT=gpurun_out/r02_u; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_rollout.py tests/test_gpu_vec_env.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log | cut -c1-300
b() { tag=$1; shift; envs=$1; shift
  env $envs timeout 400 python bench.py "$@" > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$tag.json").read().strip().splitlines()[-1])
    cl = d.get("closed_loop") or {}
    print("$tag", "%.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "closed %s" % cl.get("ms_per_step"), "launches", d["gpu_launches"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$tag failed", e, open("$OUT/bench_$tag.err").read()[-1500:])
PY
}
for r in 1 2; do
b pf1_driver FM_PREFETCH=1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3
b pf0_driver FM_PREFETCH=0 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3
done
b pf1_long FM_PREFETCH=1 --no-cpu-baseline --e2e-steps 3
b pf0_long FM_PREFETCH=0 --no-cpu-baseline --e2e-steps 3
b pf1_eager FM_PREFETCH=1 --no-cpu-baseline --e2e-steps 3 --no-step-graph
b pf0_eager FM_PREFETCH=0 --no-cpu-baseline --e2e-steps 3 --no-step-graph
b pf1_c3 FM_PREFETCH=1 --config c3 --steps 300 --warmup 25 --no-cpu-baseline --e2e-steps 3
