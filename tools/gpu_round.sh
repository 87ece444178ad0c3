#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu --set full capture of the step kernel.
# usage: tools/gpu_round.sh <tag> [kernel-regex]
set -u
TAG=${1:-rXX}
KRE=${2:-tile_kernel}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
timeout 300 python bench.py --impl reference --steps 100 --warmup 5 > $OUT/bench_reference.json 2>> $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE --launch-skip 30 -c 2 -f -o $OUT/step_kernel \
  python bench.py --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_full.log 2>&1
ls -la $OUT
