#!/bin/bash
# Round 2, GPU session K: formation walls on the device; edge-list forms; steady-state probe of the logic kernel.
set -u
OUT=gpurun_out/r02_k; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_formation.py -m gpu -q -x > $OUT/pytest_form.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_form.log
tail -25 $OUT/pytest_form.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vec_env.py -m gpu -q -x -k "edge or vec_env or spaces" > $OUT/pytest_edges.log 2>&1; tail -3 $OUT/pytest_edges.log | cut -c1-300
python tools/form_host_probe.py 2>&1 | tail -6
