#!/bin/bash
# Round 2, GPU session V: aw prefetch opt-in (default path unchanged), racecheck of the group kernels after the extra __syncwarp.
set -u
OUT=gpurun_out/r02_v; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log | cut -c1-300
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_vec_env.py tests/test_gpu_parity.py -q -x -k "soa or (prefetch_is_bitwise and 33)" > $OUT/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/sanitizer_racecheck.log | tail -3; grep -c "Race reported" $OUT/sanitizer_racecheck.log
for r in 1 2 3; do
  timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_driver$r.json 2> $OUT/bench_driver$r.err
  python -c "
import json; d=json.loads(open('$OUT/bench_driver$r.json').read().strip().splitlines()[-1]); print('driver %.4g ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done
