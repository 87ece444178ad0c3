#!/bin/bash
# same-box comparison of library builds on the config-3 step + edge-list time.  usage: tools/ab/run_edges.sh v1 v2 ...
for r in 1 2; do for v in "$@"; do
  cp tools/ab/libfairmarl_$v.so fair-marl_b200/libfairmarl.so
  python bench.py --config c3 --steps 100 --warmup 25 --no-cpu-baseline --e2e-steps 3 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['edge_list']; print('$v', d['ms_per_step'], 'with edges', e['ms_per_step_with_edge_list'], 'closed', e['ms_per_step_closed_loop'])"
done; done
