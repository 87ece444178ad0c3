#!/bin/bash
# Round 2, GPU session D: full parity suite, recorded vec-env tuples for the reference-runner test, headline (driver
# configuration and long), C3 / C4 / formation / config 5, steady-state DRAM bytes of the step kernel, launch list.
set -u
OUT=gpurun_out/r02_d; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1; nproc > $OUT/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -12 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 300 python tools/record_vec_env_tuples.py gpurun_out/r02_d/b200_vec_env_tuples.npz > $OUT/record.log 2>&1; tail -1 $OUT/record.log | cut -c1-300
b() { # tag, env assignments, args
  tag=$1; shift; envs=$1; shift
  env $envs timeout 400 python bench.py "$@" > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$tag.json").read().strip().splitlines()[-1])
    cl = d.get("closed_loop") or {}
    e2e = d.get("e2e") or {}
    print("$tag", "%.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "closed %s" % cl.get("ms_per_step"),
          "launches", d["gpu_launches"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["clocks"]["scope"], "e2e %s" % e2e.get("value"),
          "cpu", (d.get("cpu_baseline") or {}).get("value"), "edges", (d.get("edge_list") or {}).get("ms_per_step_with_edge_list"))
except Exception as e:
    print("$tag failed", e, open("$OUT/bench_$tag.err").read()[-1500:])
PY
}
b driver1 FM_X=0 --steps 20 --warmup 5
b driver2 FM_X=0 --steps 20 --warmup 5 --no-cpu-baseline
b driver3 FM_X=0 --steps 20 --warmup 5 --no-cpu-baseline
b default FM_X=0 --no-cpu-baseline
b eager_long FM_X=0 --no-cpu-baseline --no-step-graph
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2>> $OUT/bench_reference.err; cut -c1-300 $OUT/bench_reference.json
b c3 FM_X=0 --config c3 --steps 300 --warmup 25 --no-cpu-baseline --e2e-steps 3
b c4 FM_X=0 --config c4 --steps 300 --warmup 25 --no-cpu-baseline --e2e-steps 3
b c1 FM_X=0 --config c1 --steps 2000 --warmup 25 --no-cpu-baseline --e2e-steps 3
b walls_c2 FM_X=0 --walls 2 --steps 500 --warmup 25 --no-cpu-baseline --e2e-steps 3
b walls_c3 FM_X=0 --config c3 --walls 2 --steps 200 --warmup 25 --no-cpu-baseline --e2e-steps 3
b form FM_X=0 --config form --steps 300 --warmup 30
b form_short FM_X=0 --config form --steps 20 --warmup 5
for B in 4096 65536; do
  timeout 600 python bench.py --config c5 --envs $B --steps 50 > $OUT/bench_c5_$B.json 2> $OUT/bench_c5_$B.err; cut -c1-330 $OUT/bench_c5_$B.json; tail -2 $OUT/bench_c5_$B.err
done
# launch list of the driver's configuration, and steady-state DRAM bytes of the step kernel (single pass, no cache flush, no replay)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_under_ncu.log 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  -k regex:aw_kernel --launch-skip 60 -c 200 --csv --log-file $OUT/steady_dram.csv \
  python bench.py --steps 200 --warmup 25 --no-cpu-baseline --e2e-steps 3 --no-step-graph > $OUT/steady_dram.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02_d/steady_dram.csv")) if len(r) > 10]
hdr = rows[0]; vals = {}
for r in rows[1:]:
    d = dict(zip(hdr, r))
    vals.setdefault(d.get("Metric Name"), []).append((float(d.get("Metric Value", "0").replace(",", "")), d.get("Metric Unit")))
for k, v in vals.items():
    print(k, "n", len(v), "mean", sum(x for x, _ in v) / len(v), v[0][1])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:formation_kernel --launch-skip 20 -c 1 -f -o $OUT/formation_kernel \
  python bench.py --config form --steps 30 --warmup 5 > $OUT/ncu_form.log 2>&1; tail -1 $OUT/ncu_form.log
ls $OUT | wc -l
