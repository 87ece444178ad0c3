#!/bin/bash
# Round 2, GPU session B: fused GNN policy tests, rollout-kernel stagger sweep, CUDA-graph one-shot lanes, config 5, ncu.
set -u
OUT=gpurun_out/r02_b; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_policy.py tests/test_gpu_fullsize.py tests/test_gpu_rollout.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
b() { # tag, env assignments, args
  tag=$1; shift; envs=$1; shift
  env $envs timeout 300 python bench.py --no-cpu-baseline --e2e-steps 3 "$@" > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "%.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "closed %.5f" % d["closed_loop"]["ms_per_step"],
          "launches", d["gpu_launches"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["clocks"]["scope"], "eps", d["episode_stats"]["episodes"])
except Exception as e:
    print("$tag failed", e, open("$OUT/bench_$tag.err").read()[-1500:])
PY
}
for s in 0 700 1000 1400 2000 3000; do
  b long_st$s "FM_ROLL_STAGGER_NS=$s FM_ROLL_STAGGER1_NS=$s" --steps 2000 --warmup 100
done
for s in 0 1000 1400 2000; do
  b short_st$s "FM_ROLL_STAGGER_NS=$s FM_ROLL_STAGGER1_NS=$s" --steps 20 --warmup 5
done
b graph_oneshot_short FM_ROLL=0 --steps 20 --warmup 5 --graph
b graph_oneshot_short2 FM_ROLL=0 --steps 20 --warmup 5 --graph
b graph_oneshot_long FM_ROLL=0 --steps 2000 --warmup 100 --graph
b graph_roll_short FM_ROLL_STAGGER_NS=1400 --steps 20 --warmup 5 --graph
for B in 4096 65536; do
  timeout 600 python bench.py --config c5 --envs $B --steps 50 > $OUT/bench_c5_$B.json 2> $OUT/bench_c5_$B.err; cut -c1-700 $OUT/bench_c5_$B.json; tail -2 $OUT/bench_c5_$B.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aw_roll_kernel --launch-skip 4 -c 1 -f -o $OUT/roll_kernel \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_roll.log 2>&1; tail -2 $OUT/ncu_roll.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gnn_kernel --launch-skip 10 -c 1 -f -o $OUT/gnn_kernel \
  python bench.py --config c5 --envs 65536 --steps 25 --no-graph > $OUT/ncu_gnn.log 2>&1; tail -2 $OUT/ncu_gnn.log
ls -la $OUT | head -50
