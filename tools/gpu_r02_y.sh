#!/bin/bash
# Round 2, GPU session Y: same-box A/B, serial (prev) vs warp-uniform cooperative (new) reset placement in the agent-warp kernels.
set -u
OUT=gpurun_out/r02_y; mkdir -p $OUT
cp tools/ab/libfairmarl_new.so fair-marl_b200/libfairmarl.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_vec_env.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log | cut -c1-200
for r in 1 2 3; do for v in prev new; do
  cp tools/ab/libfairmarl_$v.so fair-marl_b200/libfairmarl.so
  timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > $OUT/b.json 2> $OUT/b.err
  python -c "
import json; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('$v driver %.4g ms/step %.5f frac %.3f closed %.5f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['closed_loop']['ms_per_step']))" || tail -5 $OUT/b.err
done; done
for v in prev new; do
  cp tools/ab/libfairmarl_$v.so fair-marl_b200/libfairmarl.so
  timeout 400 python bench.py --no-cpu-baseline --e2e-steps 3 --no-step-graph > $OUT/b.json 2> $OUT/b.err
  python -c "
import json; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('$v long eager %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))" || tail -5 $OUT/b.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aw_kernel --launch-skip 150 -c 40 --csv --log-file $OUT/launches_$v.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_y/launches_$v.csv")) if len(r)>10]
hdr=rows[0]; v=[float(dict(zip(hdr,r))["Metric Value"].replace(",",""))/1e3 for r in rows[1:]]
print("$v launch durations us:", " ".join("%.0f"%x for x in v))
PY
done
cp tools/ab/libfairmarl_new.so fair-marl_b200/libfairmarl.so
