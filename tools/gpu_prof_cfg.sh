#!/bin/bash
# launch list + ncu full capture of the step kernel for a given bench config.  usage: tools/gpu_prof_cfg.sh <tag> <config> <kernel-regex> <skip>
set -u
TAG=$1; CFG=$2; KRE=$3; SKIP=${4:-30}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_$CFG.csv \
  python bench.py --config $CFG --steps 60 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/under_ncu_$CFG.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/launches_$CFG.csv")))
hdr = [r for r in rows if r and r[0] == "ID"][0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
for r in rows:
    if len(r) > vi and r[0].isdigit() and "step_kernel" in r[ki] or (len(r) > vi and r[0].isdigit() and "aw_kernel" in r[ki]):
        print(r[0], r[ki][:40], r[vi])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE --launch-skip $SKIP -c 1 -f -o $OUT/step_kernel_$CFG \
  python bench.py --config $CFG --steps 60 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_full_$CFG.log 2>&1
tail -2 $OUT/ncu_full_$CFG.log
