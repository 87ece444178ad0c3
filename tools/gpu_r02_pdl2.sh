#!/bin/bash
# Round 2: programmatic dependent launches inside fm_step_many's lanes (FM_MANY_PDL=1), eager and under graph capture: A-B.
set -u
OUT=gpurun_out/${FM_OUT_TAG:-r02_pdl2}; mkdir -p $OUT
FM_MANY_PDL=1 timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x -k "rollout or lanes or prefetch or graph" > $OUT/pytest_many_pdl.log 2>&1; tail -2 $OUT/pytest_many_pdl.log | cut -c1-300
b() { tag=$1; v=$2; shift; shift
  FM_MANY_PDL=$v timeout 300 python bench.py "$@" --no-cpu-baseline --e2e-steps 3 > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "us/step %.3f" % (1e3 * d["ms_per_step"]), "frac %.3f" % d["roofline"]["frac"], d["episode_stats"])
except Exception as e:
    print("$tag failed", e, open("$OUT/bench_$tag.err").read()[-1200:])
PY
}
for r in 1 2; do
  b driver_pdl0_$r 0 --steps 20 --warmup 5
  b driver_pdl1_$r 1 --steps 20 --warmup 5
done
b long_pdl0 0
b long_pdl1 1
b long_pdl0b 0
b long_pdl1b 1
