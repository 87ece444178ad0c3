#!/bin/bash
# Round 2, GPU session E: formation kernel after the distance-matrix / emission index rework.
set -u
OUT=gpurun_out/r02_e; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_formation.py tests/test_gpu_fullsize.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log | cut -c1-300
for tag in form; do
  timeout 300 python bench.py --config form --steps 300 --warmup 30 > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python -c "
import json; d=json.loads(open('$OUT/bench_$tag.json').read().strip().splitlines()[-1]); print('$tag %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:formation_kernel --launch-skip 20 -c 1 -f -o $OUT/formation_kernel \
  python bench.py --config form --steps 30 --warmup 5 > $OUT/ncu_form.log 2>&1; tail -1 $OUT/ncu_form.log | cut -c1-200
