#!/bin/bash
# Round 2, last captures: ncu --set full of the agent-warp wall kernel and of the final edge-list emission.
set -u
OUT=gpurun_out/${FM_OUT_TAG:-r02_ncu_last}; mkdir -p $OUT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:aw_kernel --launch-skip 40 -c 1 -f -o $OUT/aw_walls \
  python bench.py --walls 2 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-step-graph > $OUT/ncu_aww.log 2>&1; tail -1 $OUT/ncu_aww.log | cut -c1-120
timeout 400 ncu --set full --clock-control none --import-source on -k regex:es_emit --launch-skip 5 -c 1 -f -o $OUT/es_emit \
  python bench.py --config c3 --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_es.log 2>&1; tail -1 $OUT/ncu_es.log | cut -c1-120
