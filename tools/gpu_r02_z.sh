#!/bin/bash
# Round 2, GPU session Z: host step as env-range lanes (fm_step_host_lane): parity and e2e A/B.
set -u
OUT=gpurun_out/r02_z; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_vec_env.py tests/test_gpu_rollout.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log | cut -c1-300
for r in 1 2; do for L in 1 2 4 8; do
  FM_HOST_LANES=$L timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 30 > $OUT/b.json 2> $OUT/b.err
  python -c "
import json; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('host lanes $L: e2e %.4g agent-steps/s, %.3f ms/step; value frac %.3f' % (d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))" || tail -5 $OUT/b.err
done; done
FM_HOST_LANES=4 timeout 400 python bench.py --config c3 --steps 100 --warmup 25 --no-cpu-baseline --e2e-steps 10 > $OUT/b.json 2> $OUT/b.err; python -c "
import json; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('c3 host lanes 4: e2e %.4g %.3f ms/step' % (d['e2e']['value'], d['e2e']['ms_per_step']))"
FM_HOST_LANES=1 timeout 400 python bench.py --config c3 --steps 100 --warmup 25 --no-cpu-baseline --e2e-steps 10 > $OUT/b.json 2> $OUT/b.err; python -c "
import json; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('c3 host lanes 1: e2e %.4g %.3f ms/step' % (d['e2e']['value'], d['e2e']['ms_per_step']))"
