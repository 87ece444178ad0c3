#!/bin/bash
# Round 2, GPU session X: CTA-cooperative reset placement in the agent-warp kernels.
set -u
OUT=gpurun_out/r02_x; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
for r in 1 2 3; do
  timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_driver$r.json 2> $OUT/bench_driver$r.err
  python -c "
import json; d=json.loads(open('$OUT/bench_driver$r.json').read().strip().splitlines()[-1]); print('driver %.4g ms/step %.5f frac %.3f closed %.5f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['closed_loop']['ms_per_step']))" || tail -5 $OUT/bench_driver$r.err
done
for a in "" "--no-step-graph" "--config c1 --steps 2000"; do
  timeout 400 python bench.py $a --no-cpu-baseline --e2e-steps 3 > $OUT/b.json 2> $OUT/b.err
  python -c "
import json; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('[$a] %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))" || tail -5 $OUT/b.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aw_kernel --launch-skip 150 -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_x/launches.csv")) if len(r)>10]
hdr=rows[0]; v=[float(dict(zip(hdr,r))["Metric Value"].replace(",",""))/1e3 for r in rows[1:]]
print("launch durations us:", " ".join("%.0f"%x for x in v))
PY
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "prefetch_is_bitwise and aw" > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/sanitizer_racecheck.log | tail -2
