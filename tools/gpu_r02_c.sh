#!/bin/bash
# Round 2, GPU session C: fused policy (GNN + head), new formation kernels (+ compute-sanitizer), rollout-kernel variants.
set -u
OUT=gpurun_out/r02_c; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_policy.py tests/test_gpu_formation.py tests/test_gpu_fullsize.py tests/test_gpu_rollout.py -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -30 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
for tool in memcheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_formation.py -q -x -k "reset_and_rollout and (4-2 or 7-3 or 3-3-70) or masked" > $OUT/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?" | tee -a $OUT/sanitizer_$tool.log; grep -E "ERROR SUMMARY|passed|failed" $OUT/sanitizer_$tool.log | tail -3
done
b() { # tag, env assignments, args
  tag=$1; shift; envs=$1; shift
  env $envs timeout 300 python bench.py --no-cpu-baseline --e2e-steps 3 "$@" > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$tag.json").read().strip().splitlines()[-1])
    cl = d.get("closed_loop") or {}
    print("$tag", "%.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "closed %s" % cl.get("ms_per_step"),
          "launches", d["gpu_launches"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["clocks"]["scope"])
except Exception as e:
    print("$tag failed", e, open("$OUT/bench_$tag.err").read()[-1500:])
PY
}
b roll_long FM_X=0 --steps 2000 --warmup 100
b roll_short FM_X=0 --steps 20 --warmup 5
b roll_short2 FM_X=0 --steps 20 --warmup 5
b roll_short_graph FM_X=0 --steps 20 --warmup 5 --graph
b oneshot_short_graph FM_ROLL=0 --steps 20 --warmup 5 --graph
b oneshot_long_graph FM_ROLL=0 --steps 2000 --warmup 100 --graph
b form FM_X=0 --config form --steps 300 --warmup 30
b form_short FM_X=0 --config form --steps 20 --warmup 5
b form_1slot FM_X=0 --config form --steps 300 --warmup 30 --form-slots 1
for B in 4096 65536; do
  timeout 600 python bench.py --config c5 --envs $B --steps 50 > $OUT/bench_c5_$B.json 2> $OUT/bench_c5_$B.err; cut -c1-330 $OUT/bench_c5_$B.json; tail -2 $OUT/bench_c5_$B.err
done
timeout 600 python bench.py --config c5 --envs 65536 --steps 50 --no-graph > $OUT/bench_c5_65536_eager.json 2> $OUT/bench_c5_eager.err; cut -c1-330 $OUT/bench_c5_65536_eager.json
timeout 300 python tools/prof_policy.py 65536 > $OUT/policy_profile.txt 2>&1; head -30 $OUT/policy_profile.txt | cut -c1-180
timeout 600 ncu --set full --clock-control none --import-source on -k regex:formation_kernel --launch-skip 20 -c 1 -f -o $OUT/formation_kernel \
  python bench.py --config form --steps 30 --warmup 5 > $OUT/ncu_form.log 2>&1; tail -2 $OUT/ncu_form.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_kernel --launch-skip 10 -c 1 -f -o $OUT/head_kernel \
  python bench.py --config c5 --envs 65536 --steps 25 --no-graph > $OUT/ncu_head.log 2>&1; tail -2 $OUT/ncu_head.log
ls $OUT | head -60
