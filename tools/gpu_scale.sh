#!/bin/bash
# headline line at N GPUs (torchrun).  usage: tools/gpu_scale.sh <ngpus>
N=$1; mkdir -p gpurun_out/scale
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5000 --warmup 100 --no-cpu-baseline > gpurun_out/scale/bench_c2_${N}gpu.json 2> gpurun_out/scale/err_$N.txt
tail -1 gpurun_out/scale/bench_c2_${N}gpu.json | cut -c1-220
