#!/bin/bash
# Round 2, GPU session M: formation with the stage-less logic tile + max carve-out; carve-out A/B on the navigation kernels.
set -u
OUT=gpurun_out/r02_m; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_formation.py -m gpu -q -x > $OUT/pytest_form.log 2>&1; tail -2 $OUT/pytest_form.log
python tools/form_host_probe.py 2>&1 | tail -9 | cut -c1-420
for r in 1 2; do
  timeout 300 python bench.py --config form --steps 300 --warmup 30 > $OUT/bench_form.json 2> $OUT/bench_form.err
  python -c "
import json; d=json.loads(open('$OUT/bench_form.json').read().strip().splitlines()[-1]); print('form %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done
for cv in 0 1 0 1; do
 for c in "" "--config c3 --steps 200" "--config c4 --steps 200" "--walls 2 --steps 300"; do
  FM_CARVEOUT=$cv timeout 300 python bench.py $c --warmup 25 --no-cpu-baseline --e2e-steps 3 > $OUT/b.json 2> $OUT/b.err
  python -c "
import json; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('carveout=$cv [$c] ms/step %.5f frac %.3f closed %s' % (d['ms_per_step'], d['roofline']['frac'], d['closed_loop']['ms_per_step']))"
 done
done
