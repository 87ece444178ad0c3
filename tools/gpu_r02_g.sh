#!/bin/bash
# Round 2, GPU session G: formation split path (logic kernel + image kernel) against the fused kernel.
set -u
OUT=gpurun_out/r02_g; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_formation.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log | cut -c1-300
for r in 1 2; do for v in split fused; do
  FUSED=0; [ $v = fused ] && FUSED=1
  FM_FORM_FUSED=$FUSED timeout 300 python bench.py --config form --steps 300 --warmup 30 > $OUT/bench_$v.json 2> $OUT/bench_$v.err
  python -c "
import json; d=json.loads(open('$OUT/bench_$v.json').read().strip().splitlines()[-1]); print('$v %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:formation --launch-skip 40 -c 8 --csv --log-file $OUT/form_launches.csv python bench.py --config form --steps 30 --warmup 5 > /dev/null 2>&1; grep -v "^==" $OUT/form_launches.csv | cut -d, -f5,15- | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:formation --launch-skip 40 -c 2 -f -o $OUT/formation_split \
  python bench.py --config form --steps 30 --warmup 5 > $OUT/ncu_form.log 2>&1; tail -1 $OUT/ncu_form.log | cut -c1-200
