#!/bin/bash
# Multi-GPU session: host D2H ceiling at 1/2/4/N ranks, then the headline bench at N ranks (driver configuration and long).
# usage (gpurun --gpus N): bash tools/gpu_r02_multi.sh N
set -u
N=${1:-8}; OUT=gpurun_out/r02_multi; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc > $OUT/nproc.txt; lscpu | grep -E "NUMA|Socket|Model name" > $OUT/lscpu.txt 2>&1
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 tools/pcie_probe_multi.py 2> $OUT/probe_$n.err | tail -1 | tee -a $OUT/pcie_probe.jsonl | cut -c1-400
done
for n in 2 8; do
  [ $n -le $N ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_driver_${n}gpu.json 2> $OUT/bench_driver_${n}gpu.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_driver_${n}gpu.json").read().strip().splitlines()[-1])
    print("$n gpus", "%.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"], "eps", d["episode_stats"])
except Exception as e:
    print("$n failed", e, open("$OUT/bench_driver_${n}gpu.err").read()[-1500:])
PY
done
ls $OUT
