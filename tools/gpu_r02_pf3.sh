#!/bin/bash
# Round 2: sliced next-episode prefetch by the env warp of the agent-warp step kernel (FM_PREFETCH=3): parity, A-B, launch list.
set -u
OUT=gpurun_out/${FM_OUT_TAG:-r02_pf3}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "prefetch" > $OUT/pytest_prefetch.log 2>&1; tail -3 $OUT/pytest_prefetch.log | cut -c1-300
FM_PREFETCH=3 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vec_env.py tests/test_gpu_fullsize.py tests/test_gpu_rollout.py -m gpu -q -x > $OUT/pytest_all_pf3.log 2>&1; tail -3 $OUT/pytest_all_pf3.log | cut -c1-300
b() { # tag, pf, args
  tag=$1; pf=$2; shift; shift
  FM_PREFETCH=$pf timeout 400 python bench.py "$@" > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "%.4g" % d["value"], "us/step %.3f" % (1e3 * d["ms_per_step"]), "frac %.3f" % d["roofline"]["frac"], "closed %.3f" % (1e3 * d["closed_loop"]["ms_per_step"]), "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print("$tag failed", e, open("$OUT/bench_$tag.err").read()[-1500:])
PY
}
for r in 1 2; do
  b driver_pf1_$r 1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3
  b driver_pf3_$r 3 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3
done
b long_pf1 1 --no-cpu-baseline --e2e-steps 3
b long_pf3 3 --no-cpu-baseline --e2e-steps 3
b long_pf1b 1 --no-cpu-baseline --e2e-steps 3
b long_pf3b 3 --no-cpu-baseline --e2e-steps 3
for pf in 1 3; do
FM_PREFETCH=$pf timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aw_kernel -c 120 --csv --log-file $OUT/launches_pf$pf.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-step-graph > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(l for l in open("$OUT/launches_pf$pf.csv") if not l.startswith("=="))]
t = [float(r[-1]) / 1e3 for r in rows[1:] if "aw_kernel<3, 3, 0" in r[4]]
print("pf$pf launches", len(t), "median %.1f" % sorted(t)[len(t) // 2], "max %.1f" % max(t), "top", sorted(t)[-6:])
PY
done
