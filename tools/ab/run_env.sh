#!/bin/bash
# same-box comparison of FM_STAGE variants.  usage: tools/ab/run_env.sh <config> v1 v2 ...   (v = default | k1x2 | k3x1 | k3x2 | k4x1)
C=$1; shift
for r in 1 2; do for v in "$@"; do
  if [ $v = default ]; then unset FM_STAGE; else export FM_STAGE=$v; fi
  python bench.py --config $C --steps 300 --warmup 25 --no-cpu-baseline --e2e-steps 3 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$C', '$v', d['ms_per_step'], d['roofline']['frac'])"
done; done
