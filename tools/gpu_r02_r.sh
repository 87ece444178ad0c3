#!/bin/bash
# Round 2, GPU session R: fast reset tail (pending blocks + unrolled re-observation) with the dependent launch kept.
set -u
OUT=gpurun_out/r02_r; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_formation.py -m gpu -q -x > $OUT/pytest_form.log 2>&1; tail -4 $OUT/pytest_form.log | cut -c1-300
FM_FORM_PREFETCH=0 timeout 600 python -m pytest tests/test_gpu_formation.py -m gpu -q -x -k "rollout" > $OUT/pytest_form_nopf.log 2>&1; tail -1 $OUT/pytest_form_nopf.log
FM_FORM_PDL=0 timeout 600 python -m pytest tests/test_gpu_formation.py -m gpu -q -x -k "rollout" > $OUT/pytest_form_nopdl.log 2>&1; tail -1 $OUT/pytest_form_nopdl.log
FM_FORM_PREFETCH=1 python tools/form_tail_probe.py 2>&1 | head -1 | cut -c1-330
for r in 1 2; do for pf in 1 0; do
  FM_FORM_PREFETCH=$pf timeout 300 python bench.py --config form --steps 300 --warmup 30 > $OUT/bench_form_pf$pf.json 2> $OUT/bench_form_pf$pf.err
  python -c "
import json; d=json.loads(open('$OUT/bench_form_pf$pf.json').read().strip().splitlines()[-1]); print('form prefetch=$pf %.4g ms/step %.5f frac %.3f closed %.5f (%.3f)' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['closed_loop']['ms_per_step'], d['closed_loop']['frac']))" || tail -5 $OUT/bench_form_pf$pf.err
done; done
timeout 300 python bench.py --config form --steps 20 --warmup 5 > $OUT/bench_form_short.json 2> $OUT/bench_form_short.err; python -c "
import json; d=json.loads(open('$OUT/bench_form_short.json').read().strip().splitlines()[-1]); print('form short %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
