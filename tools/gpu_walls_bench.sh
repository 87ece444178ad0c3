#!/bin/bash
# Wall instantiations measured: bench lines at C2 / C3 shapes with 2 walls, launch list and one full ncu capture of
# step_kernel<8, true>.  usage: tools/gpu_walls_bench.sh <tag>
set -u
OUT=gpurun_out/${1:-walls_bench}; mkdir -p $OUT
for cfg in c2 c3; do
  steps=1000; [ $cfg = c3 ] && steps=300
  timeout 80 python bench.py --config $cfg --walls 2 --steps $steps --warmup 50 --no-cpu-baseline --e2e-steps 3 2> $OUT/bench_${cfg}_w2.err | tail -1 > $OUT/bench_${cfg}_w2.json
  cut -c1-600 $OUT/bench_${cfg}_w2.json; tail -3 $OUT/bench_${cfg}_w2.err
done
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_c3_w2.csv \
  python bench.py --config c3 --walls 2 --steps 60 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/under_ncu_c3_w2.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 30 -c 1 -f -o $OUT/step_kernel_c3_w2 \
  python bench.py --config c3 --walls 2 --steps 60 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_full_c3_w2.log 2>&1
ls -la $OUT
