#!/bin/bash
# Round 2, GPU session S: env-range lanes again, now that the reset tail is short.
set -u
OUT=gpurun_out/r02_s; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_formation.py -m gpu -q -x -k "lanes" > $OUT/pytest_form.log 2>&1; tail -2 $OUT/pytest_form.log | cut -c1-300
for r in 1 2 3; do for lanes in 2 1; do
  FM_FORM_LANES=$lanes timeout 300 python bench.py --config form --steps 300 --warmup 30 > $OUT/bench_form_l$lanes.json 2> $OUT/bench_form_l$lanes.err
  python -c "
import json; d=json.loads(open('$OUT/bench_form_l$lanes.json').read().strip().splitlines()[-1]); print('form lanes=$lanes %.4g ms/step %.5f frac %.3f closed %.5f (%.3f)' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['closed_loop']['ms_per_step'], d['closed_loop']['frac']))" || tail -5 $OUT/bench_form_l$lanes.err
done; done
