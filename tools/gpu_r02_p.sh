#!/bin/bash
# Round 2, GPU session P: logic -> image as a programmatic dependent launch with per-16-env ready flags.
set -u
OUT=gpurun_out/r02_p; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_formation.py -m gpu -q -x > $OUT/pytest_form.log 2>&1; tail -3 $OUT/pytest_form.log | cut -c1-300
FM_FORM_PREFETCH=1 timeout 600 python -m pytest tests/test_gpu_formation.py -m gpu -q -x -k "rollout" > $OUT/pytest_form_pf.log 2>&1; tail -1 $OUT/pytest_form_pf.log
FM_FORM_PDL=0 timeout 600 python -m pytest tests/test_gpu_formation.py -m gpu -q -x -k "rollout" > $OUT/pytest_form_nopdl.log 2>&1; tail -1 $OUT/pytest_form_nopdl.log
python tools/form_host_probe.py 2>&1 | sed -n 1,6p | cut -c1-200
for r in 1 2; do for pdl in 1 0; do
  FM_FORM_PDL=$pdl timeout 300 python bench.py --config form --steps 300 --warmup 30 > $OUT/bench_form_pdl$pdl.json 2> $OUT/bench_form_pdl$pdl.err
  python -c "
import json; d=json.loads(open('$OUT/bench_form_pdl$pdl.json').read().strip().splitlines()[-1]); print('form pdl=$pdl %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done; done
timeout 300 python bench.py --config form --steps 20 --warmup 5 > $OUT/bench_form_short.json 2> $OUT/bench_form_short.err; python -c "
import json; d=json.loads(open('$OUT/bench_form_short.json').read().strip().splitlines()[-1]); print('form short %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
