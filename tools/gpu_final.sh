#!/bin/bash
# Round artefacts in one GPU-box pass: parity suite, headline bench (+ reference arm), diagnostic configs, rollout loop,
# ncu launch lists and full captures (headline kernel, C3 / C4 group kernel).  usage: tools/gpu_final.sh <tag>
set -u
TAG=$1; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1; nproc > $OUT/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json | cut -c1-400
timeout 300 python bench.py --impl reference --steps 100 --warmup 5 > $OUT/bench_reference.json 2>> $OUT/bench.err
for c in c1 c3 c4; do
  steps=300; [ $c = c1 ] && steps=5000
  timeout 600 python bench.py --config $c --steps $steps --warmup 25 --e2e-steps 5 > $OUT/bench_$c.json 2> $OUT/bench_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$c.json").read().strip().splitlines()[-1])
    print("$c", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, "edges", d.get("edge_list"))
except Exception as e:
    print("$c failed", e, open("$OUT/bench_$c.err").read()[-1500:])
PY
done
for B in 4096 65536; do
  timeout 600 python bench.py --config c5 --envs $B --steps 50 > $OUT/bench_c5_$B.json 2> $OUT/bench_c5_$B.err; cut -c1-300 $OUT/bench_c5_$B.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aw_kernel --launch-skip 30 -c 2 -f -o $OUT/step_kernel \
  python bench.py --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_full.log 2>&1
for c in c3 c4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_$c.csv \
    python bench.py --config $c --steps 60 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/under_ncu_$c.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 30 -c 1 -f -o $OUT/step_kernel_$c \
    python bench.py --config $c --steps 60 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_full_$c.log 2>&1
done
ls -la $OUT
