#!/bin/bash
# Round 2: programmatic dependent launch between consecutive single-step launches of the agent-warp kernel (FM_STEP_PDL=1):
# parity of the single-step API, closed-loop A-B (the headline rollout path does not use the attribute).
set -u
OUT=gpurun_out/${FM_OUT_TAG:-r02_pdl}; mkdir -p $OUT
FM_STEP_PDL=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vec_env.py tests/test_gpu_rollout.py -m gpu -q -x -k "not edge" > $OUT/pytest_pdl.log 2>&1; tail -2 $OUT/pytest_pdl.log | cut -c1-200
for v in 0 1 0 1; do
  FM_STEP_PDL=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_pdl$v.json 2> $OUT/bench_pdl$v.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_pdl$v.json").read().strip().splitlines()[-1])
    print("pdl$v", "us/step %.3f" % (1e3 * d["ms_per_step"]), "frac %.3f" % d["roofline"]["frac"], "closed %.3f us" % (1e3 * d["closed_loop"]["ms_per_step"]), "= %.3f of the roofline" % (d["roofline"]["algorithmic_bytes_per_step"] / (d["closed_loop"]["ms_per_step"] * 1e-3) / 1e9 / d["roofline"]["peak"]))
except Exception as e:
    print("pdl$v failed", e, open("$OUT/bench_pdl$v.err").read()[-1500:])
PY
done
for v in 0 1; do
  FM_STEP_PDL=$v timeout 300 python bench.py --config c5 --envs 65536 --steps 50 > $OUT/bench_c5_pdl$v.json 2> $OUT/bench_c5_pdl$v.err; cut -c1-230 $OUT/bench_c5_pdl$v.json | cut -c100-230; echo
done
