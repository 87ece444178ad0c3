#!/bin/bash
# Quick GPU pass: selected tests + short bench variants.  usage: tools/gpu_quick.sh <tag> "<pytest -k expr>" [bench args...]
set -u
TAG=$1; KEXPR=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -15 $OUT/pytest.log
for variant in "$@"; do
  echo "== bench $variant"
  timeout 300 python bench.py --steps 2000 --warmup 50 --no-cpu-baseline --e2e-steps 5 $variant 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l); print(json.dumps({k: d[k] for k in ('value','ms_per_step')}), d['roofline']['frac'], d['roofline']['kernel'], d['e2e']['value'])
except Exception as e:
    print('bench failed:', l[-2000:])
" | tee -a $OUT/bench_variants.txt
done
