#!/bin/bash
# Round-2 first step for the open N = 4 finding of DESIGN.md section 9 (formation kernel): run the formation tests of
# the tree as it is (re-land commit 8626b1a with `git revert 69accb0` first) under compute-sanitizer, optimised and -G.
# usage: tools/gpu_formation_debug.sh <tag>
set -u
OUT=gpurun_out/${1:-form_debug}; mkdir -p $OUT
K='n4 or 4-2'
timeout 120 python -m pytest tests/test_gpu_formation.py -q -k "$K" > $OUT/plain.log 2>&1; echo "plain rc=$?"; tail -3 $OUT/plain.log
for tool in memcheck initcheck; do
  timeout 400 compute-sanitizer --tool $tool --log-file $OUT/$tool.log python -m pytest tests/test_gpu_formation.py -q -x -k "$K" > $OUT/${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; grep -c "ERROR SUMMARY\|Invalid\|Uninitialized" $OUT/$tool.log; tail -3 $OUT/$tool.log
done
FM_NVCC_EXTRA="-G" python -c "import fair_marl_b200 as f; f.build_library(force=True)" > $OUT/build_G.log 2>&1
timeout 300 python -m pytest tests/test_gpu_formation.py -q -k "$K" > $OUT/debug_build.log 2>&1; echo "-G build rc=$?"; tail -3 $OUT/debug_build.log
python -c "import fair_marl_b200 as f; f.build_library(force=True)" >> $OUT/build_G.log 2>&1     # back to the optimised library
