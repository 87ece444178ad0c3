#!/bin/bash
# Round 2: 8-GPU run of the driver's launch line with the final code (value, e2e with host lanes), and the formation line.
set -u
N=${1:-8}; OUT=gpurun_out/r02_8gpu; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err
python -c "
import json; d=json.loads(open('$OUT/bench_${N}gpu.json').read().strip().splitlines()[-1]); print('$N gpus %.4g ms/step %.5f frac %.3f e2e %.4g (%.2f ms) eps %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['episode_stats']))" || tail -20 $OUT/bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N > $OUT/bench_long_${N}gpu.json 2> $OUT/bench_long_${N}gpu.err
python -c "
import json; d=json.loads(open('$OUT/bench_long_${N}gpu.json').read().strip().splitlines()[-1]); print('$N gpus long %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))" || tail -20 $OUT/bench_long_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $N --config form --steps 300 --warmup 30 > $OUT/bench_form_${N}gpu.json 2> $OUT/bench_form_${N}gpu.err
python -c "
import json; d=json.loads(open('$OUT/bench_form_${N}gpu.json').read().strip().splitlines()[-1]); print('form $N gpus %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))" || tail -20 $OUT/bench_form_${N}gpu.err
