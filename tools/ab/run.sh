#!/bin/bash
# A/B on one box: alternate two builds of libfairmarl.so.  usage: tools/ab/run.sh "<bench args>" [rounds]
ARGS=$1; R=${2:-2}
for r in $(seq $R); do for v in prev new; do
  cp tools/ab/libfairmarl_$v.so fair-marl_b200/libfairmarl.so
  python bench.py $ARGS --no-cpu-baseline --e2e-steps 3 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '$ARGS', d['ms_per_step'], d['roofline']['frac'])"
done; done
cp tools/ab/libfairmarl_new.so fair-marl_b200/libfairmarl.so
