#!/bin/bash
# Round 2, final GPU session: full parity suite, smoke, every bench line in the driver's configuration, ncu launch list +
# full captures of the headline and formation kernels, steady-state DRAM bytes, compute-sanitizer over the new paths.
set -u
OUT=gpurun_out/${FM_OUT_TAG:-r02_final}; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1; nproc > $OUT/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
b() { # tag, args
  tag=$1; shift
  timeout 400 python bench.py "$@" > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$tag.json").read().strip().splitlines()[-1])
    cl = d.get("closed_loop") or {}
    e2e = d.get("e2e") or {}
    print("$tag", "%.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "closed %s" % cl.get("ms_per_step"),
          "launches", d["gpu_launches"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["clocks"]["scope"], "e2e %s" % e2e.get("value"),
          "cpu", (d.get("cpu_baseline") or {}).get("value"), "edges", (d.get("edge_list") or {}).get("ms_per_step_with_edge_list"), "eps", d.get("episode_stats"))
except Exception as e:
    print("$tag failed", e, open("$OUT/bench_$tag.err").read()[-1500:])
PY
}
b driver1 --steps 20 --warmup 5
b driver2 --steps 20 --warmup 5 --no-cpu-baseline
b driver3 --steps 20 --warmup 5 --no-cpu-baseline
b default --no-cpu-baseline
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2>> $OUT/bench_reference.err; cut -c1-260 $OUT/bench_reference.json
b c1 --config c1 --steps 2000 --warmup 25 --no-cpu-baseline --e2e-steps 3
b c3 --config c3 --steps 300 --warmup 25 --no-cpu-baseline --e2e-steps 3
b c4 --config c4 --steps 300 --warmup 25 --no-cpu-baseline --e2e-steps 3
b walls_c2 --walls 2 --steps 500 --warmup 25 --no-cpu-baseline --e2e-steps 3
b form --config form --steps 300 --warmup 30
b form_short --config form --steps 20 --warmup 5
for B in 4096 65536; do
  timeout 600 python bench.py --config c5 --envs $B --steps 50 > $OUT/bench_c5_$B.json 2> $OUT/bench_c5_$B.err; cut -c1-200 $OUT/bench_c5_$B.json; echo
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:formation --launch-skip 30 -c 60 --csv --log-file $OUT/form_launches.csv \
  python bench.py --config form --steps 60 --warmup 5 > /dev/null 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  -k regex:formation --launch-skip 60 -c 100 --csv --log-file $OUT/form_steady_dram.csv python bench.py --config form --steps 100 --warmup 25 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:formation_ --launch-skip 40 -c 2 -f -o $OUT/formation_final \
  python bench.py --config form --steps 30 --warmup 5 > $OUT/ncu_form.log 2>&1; tail -1 $OUT/ncu_form.log | cut -c1-160
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aw_kernel --launch-skip 40 -c 1 -f -o $OUT/aw_kernel_final \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-step-graph > $OUT/ncu_aw.log 2>&1; tail -1 $OUT/ncu_aw.log | cut -c1-160
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_formation.py tests/test_gpu_vec_env.py tests/test_gpu_parity.py -q -x \
    -k "(rollout_lanes and 300) or (reset_and_rollout and (3-2-40 or 3-3-70 or 4-2-32)) or soa or finite_guard or edge_list_corner" > $OUT/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?" | tee -a $OUT/sanitizer_$tool.log; grep -E "ERROR SUMMARY|passed|failed" $OUT/sanitizer_$tool.log | tail -3
done
ls $OUT | wc -l
