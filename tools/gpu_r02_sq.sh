#!/bin/bash
# Round 2: reset placement of the agent-warp kernels (squared-distance tests, register-resident candidates): parity and same-box A/B.
set -u
OUT=gpurun_out/r02_sq; mkdir -p $OUT
cp tools/ab/libfairmarl_new.so fair-marl_b200/libfairmarl.so
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_vec_env.py tests/test_gpu_rollout.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log | cut -c1-200
bash tools/gpu_r02_ab.sh "--steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3" prev new
bash tools/gpu_r02_ab.sh "--no-cpu-baseline --e2e-steps 3 --no-step-graph" prev new 2>&1 | head -4
cp tools/ab/libfairmarl_new.so fair-marl_b200/libfairmarl.so
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aw_kernel --launch-skip 150 -c 40 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_sq/launches.csv")) if len(r)>10]
hdr=rows[0]; v=[float(dict(zip(hdr,r))["Metric Value"].replace(",",""))/1e3 for r in rows[1:]]
print("launch durations us:", " ".join("%.0f"%x for x in v))
PY
