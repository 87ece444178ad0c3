#!/bin/bash
# ncu --set full capture of the step kernel under a short bench run.  usage: tools/gpu_prof.sh <tag> <kernel-regex> [bench args]
set -u
TAG=$1; KRE=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE --launch-skip 30 -c 2 -f -o $OUT/step_kernel \
  python bench.py --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 3 "$@" > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
