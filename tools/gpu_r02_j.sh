#!/bin/bash
set -u
OUT=gpurun_out/r02_j; mkdir -p $OUT
python tools/form_host_probe.py 2>&1 | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_fused --launch-skip 5 -c 1 -f -o $OUT/edge_fused \
  python bench.py --config c3 --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_edge.log 2>&1; tail -1 $OUT/ncu_edge.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:edge --launch-skip 5 -c 4 --csv --log-file $OUT/edge_launches.csv python bench.py --config c3 --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 3 > /dev/null 2>&1; grep -v "^==" $OUT/edge_launches.csv | cut -d, -f5,15- | tail -4
FM_EDGE_FUSED=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:edge --launch-skip 15 -c 6 --csv --log-file $OUT/edge3_launches.csv python bench.py --config c3 --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 3 > /dev/null 2>&1; grep -v "^==" $OUT/edge3_launches.csv | cut -d, -f5,15- | tail -6
