#!/bin/bash
# Round 2, GPU session A: full parity suite with the persistent rollout kernel + float64 statistics, the driver's
# short bench configuration, the long one, FM_ROLL / FM_ROLL_CTAS variants, fairness-error distribution.
set -u
OUT=gpurun_out/r02_a; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1; nproc > $OUT/nproc.txt
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
b() { # tag, env assignments, args
  tag=$1; shift; envs=$1; shift
  env $envs timeout 300 python bench.py --no-cpu-baseline --e2e-steps 5 "$@" > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "%.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "closed %.5f" % d["closed_loop"]["ms_per_step"],
          "launches", d["gpu_launches"], "clk", d["clocks"], "eps", d["episode_stats"])
except Exception as e:
    print("$tag failed", e, open("$OUT/bench_$tag.err").read()[-1500:])
PY
}
b short1 FM_X=0 --steps 20 --warmup 5
b short2 FM_X=0 --steps 20 --warmup 5
b short3 FM_X=0 --steps 20 --warmup 5
b long FM_X=0 --steps 5000 --warmup 100
b long_oneshot FM_ROLL=0 --steps 5000 --warmup 100
b short_oneshot FM_ROLL=0 --steps 20 --warmup 5
b long_592 FM_ROLL_CTAS=592 --steps 2000 --warmup 100
b long_444 FM_ROLL_CTAS=444 --steps 2000 --warmup 100
b long_1480 FM_ROLL_CTAS=1480 --steps 2000 --warmup 100
b c3 FM_X=0 --config c3 --steps 300 --warmup 25
b c4 FM_X=0 --config c4 --steps 300 --warmup 25
b c1 FM_X=0 --config c1 --steps 2000 --warmup 25
timeout 600 python tools/fairness_error.py > $OUT/fairness_error.jsonl 2> $OUT/fairness_error.err; cat $OUT/fairness_error.jsonl; tail -3 $OUT/fairness_error.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
ls $OUT
