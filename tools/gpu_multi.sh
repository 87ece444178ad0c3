#!/bin/bash
# N-GPU pass (torchrun, one rank per GPU): headline C2, C4 (131072 envs per GPU = 1M over 8) and the config-5 rollout loop.
# usage: tools/gpu_multi.sh <tag> <ngpus>
set -u
TAG=$1; N=$2; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 5000 --warmup 100 --no-cpu-baseline > $OUT/bench_c2_${N}gpu.json 2> $OUT/bench_c2_${N}gpu.err; tail -1 $OUT/bench_c2_${N}gpu.json | cut -c1-200
timeout 600 $TR bench.py --gpus $N --config c4 --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline > $OUT/bench_c4_${N}gpu.json 2> $OUT/bench_c4_${N}gpu.err; tail -1 $OUT/bench_c4_${N}gpu.json | cut -c1-200
timeout 600 $TR bench.py --gpus $N --config c3 --steps 300 --warmup 25 --e2e-steps 3 --no-cpu-baseline > $OUT/bench_c3_${N}gpu.json 2> $OUT/bench_c3_${N}gpu.err; tail -1 $OUT/bench_c3_${N}gpu.json | cut -c1-200
timeout 900 $TR bench.py --gpus $N --config c5 --envs 524288 --steps 25 > $OUT/bench_c5_${N}gpu_524288.json 2> $OUT/bench_c5_${N}gpu_524288.err; tail -1 $OUT/bench_c5_${N}gpu_524288.json | cut -c1-300
tail -3 $OUT/*.err | cut -c1-300
