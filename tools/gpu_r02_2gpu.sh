#!/bin/bash
# Round 2: 2-GPU sanity of the driver's launch lines (ours + reference arm) with the final code.
set -u
OUT=gpurun_out/${FM_OUT_TAG:-r02_2gpu}; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
python -c "
import json; d=json.loads(open('$OUT/bench_2gpu.json').read().strip().splitlines()[-1]); print('2 gpus %.4g ms/step %.5f frac %.3f e2e %.4g eps %s cpu %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['episode_stats'], (d.get('cpu_baseline') or {}).get('value')))" || tail -20 $OUT/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $OUT/bench_ref_2gpu.json 2> $OUT/bench_ref_2gpu.err; cut -c1-200 $OUT/bench_ref_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --config form --steps 100 --warmup 10 > $OUT/bench_form_2gpu.json 2> $OUT/bench_form_2gpu.err
python -c "
import json; d=json.loads(open('$OUT/bench_form_2gpu.json').read().strip().splitlines()[-1]); print('form 2 gpus %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))" || tail -20 $OUT/bench_form_2gpu.err
