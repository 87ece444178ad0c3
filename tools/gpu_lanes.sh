#!/bin/bash
for L in 1 2 3 4 6 8; do
  echo "== lanes $L"
  FM_LANES=$L timeout 300 python bench.py --steps 3000 --warmup 50 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])"
done
for L in 1 4 8; do
  echo "== c3 lanes $L"; FM_LANES=$L timeout 300 python bench.py --config c3 --steps 300 --warmup 25 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])"
done
