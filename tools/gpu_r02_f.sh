#!/bin/bash
# Round 2, GPU session F: registerised formation kernel; A/B over the register cap (resident warps).
set -u
OUT=gpurun_out/r02_f; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_formation.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log | cut -c1-300
for r in 1 2; do for v in mb1 mb12 mb16; do
  cp tools/ab/libfairmarl_$v.so fair-marl_b200/libfairmarl.so
  timeout 300 python bench.py --config form --steps 300 --warmup 30 > $OUT/bench_$v.json 2> $OUT/bench_$v.err
  python -c "
import json; d=json.loads(open('$OUT/bench_$v.json').read().strip().splitlines()[-1]); print('$v %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done; done
cp tools/ab/libfairmarl_mb12.so fair-marl_b200/libfairmarl.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:formation_kernel --launch-skip 20 -c 1 -f -o $OUT/formation_kernel \
  python bench.py --config form --steps 30 --warmup 5 > $OUT/ncu_form.log 2>&1; tail -1 $OUT/ncu_form.log | cut -c1-200
