#!/bin/bash
# same-box A/B over builds of libfairmarl.so: tools/gpu_r02_ab.sh "<bench args>" v1 v2 ...
set -u
ARGS=$1; shift
for r in 1 2 3; do for v in "$@"; do
  cp tools/ab/libfairmarl_$v.so fair-marl_b200/libfairmarl.so
  timeout 400 python bench.py $ARGS > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); cl=d.get('closed_loop') or {}; print('$v [$ARGS] %.4g ms/step %.5f frac %.3f closed %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], cl.get('ms_per_step')))" || tail -5 gpurun_out/ab.err
done; done
