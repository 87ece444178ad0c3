#!/bin/bash
# ncu full capture of a RESET step of the group kernel (the 25th step of an episode).  usage: gpu_prof_reset.sh <tag> <config>
set -u
TAG=$1; CFG=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 48 -c 1 -f -o $OUT/reset_step_$CFG \
  python bench.py --config $CFG --steps 60 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_reset_$CFG.log 2>&1
tail -2 $OUT/ncu_reset_$CFG.log
