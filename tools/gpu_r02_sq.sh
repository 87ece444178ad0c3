#!/bin/bash
# Round 2: squared-distance rejection tests in the reset placement: parity (bit-exact resets) and same-box A/B.
set -u
OUT=gpurun_out/r02_sq; mkdir -p $OUT
cp tools/ab/libfairmarl_new.so fair-marl_b200/libfairmarl.so
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_vec_env.py tests/test_gpu_rollout.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log | cut -c1-200
bash tools/gpu_r02_ab.sh "--steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3" prev new
bash tools/gpu_r02_ab.sh "--no-cpu-baseline --e2e-steps 3 --no-step-graph" prev new 2>&1 | head -4
bash tools/gpu_r02_ab.sh "--config c3 --steps 300 --warmup 25 --no-cpu-baseline --e2e-steps 3" prev new 2>&1 | head -4
cp tools/ab/libfairmarl_new.so fair-marl_b200/libfairmarl.so
