#!/bin/bash
# Round 2, GPU session I: 16-env image CTAs; single-pass edge list against the three-kernel form.
set -u
OUT=gpurun_out/r02_i; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_formation.py tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log | cut -c1-300
FM_EDGE_FUSED=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k edge > $OUT/pytest_edges_3pass.log 2>&1; tail -2 $OUT/pytest_edges_3pass.log
for r in 1 2; do
  timeout 300 python bench.py --config form --steps 300 --warmup 30 > $OUT/bench_form.json 2> $OUT/bench_form.err
  python -c "
import json; d=json.loads(open('$OUT/bench_form.json').read().strip().splitlines()[-1]); print('form %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done
for v in 1 0 1 0; do
  FM_EDGE_FUSED=$v timeout 300 python bench.py --config c3 --steps 100 --warmup 25 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_c3_fused$v.json 2> $OUT/bench_c3_fused$v.err
  python -c "
import json; d=json.loads(open('$OUT/bench_c3_fused$v.json').read().strip().splitlines()[-1]); print('c3 fused=$v ms/step %.5f closed %s edges %s' % (d['ms_per_step'], d['closed_loop']['ms_per_step'], d.get('edge_list')))" | cut -c1-400
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:formation --launch-skip 40 -c 6 --csv --log-file $OUT/form_launches.csv python bench.py --config form --steps 30 --warmup 5 > /dev/null 2>&1; grep -v "^==" $OUT/form_launches.csv | cut -d, -f5,15- | tail -6
