#!/bin/bash
# Wall kernels on the GPU: wall parity tests, the golden-state tests of every config, short C2 / C3 bench lines
# (regression check of the wall-free instantiations).  usage: tools/gpu_walls.sh <tag>
set -u
OUT=gpurun_out/${1:-walls}; mkdir -p $OUT
timeout 150 python -m pytest tests -m gpu -q -k "walls or w2 or w1 or golden_states" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -25 $OUT/pytest.log
for cfg in c2 c3; do
  timeout 70 python bench.py --config $cfg --steps 1000 --warmup 50 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 > $OUT/bench_$cfg.json
  python - $OUT/bench_$cfg.json <<'PY'
import sys, json
l = open(sys.argv[1]).read().strip()
try:
    d = json.loads(l); print(d['config']['workload'][:40], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
except Exception as e:
    print('bench failed:', l[-1500:])
PY
done
