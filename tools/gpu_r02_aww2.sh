#!/bin/bash
# Round 2: compute-sanitizer over the agent-warp wall kernels, smoke of the final tree.
set -u
OUT=gpurun_out/${FM_OUT_TAG:-r02_aww2}; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -7 $OUT/smoke.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x \
    -k "(mappings_bitwise and (3-3-333-1 or 4-2-200-1)) or (walls_reset and (3-3-1-70 or 3-3-2-192) and auto)" > $OUT/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?" | tee -a $OUT/sanitizer_$tool.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/sanitizer_$tool.log | tail -3
done
