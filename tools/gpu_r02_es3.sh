#!/bin/bash
# Round 2: streamed edge list, emission at 5 CTAs / SM (48 registers): parity + config 3 timing + launch list.
set -u
OUT=gpurun_out/${FM_OUT_TAG:-r02_es3}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vec_env.py -m gpu -q -x -k "edge or adjacency" > $OUT/pytest_edges.log 2>&1; tail -3 $OUT/pytest_edges.log | cut -c1-300
for v in stream three stream; do
  FM_EDGE_FORM=$v timeout 300 python bench.py --config c3 --steps 100 --warmup 25 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_c3_$v.json 2> $OUT/bench_c3_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_c3_$v.json").read().strip().splitlines()[-1])
    print("$v", "step %.4f" % d["ms_per_step"], "closed %.4f" % d["closed_loop"]["ms_per_step"], "with edges %.4f" % d["edge_list"]["ms_per_step_with_edge_list"], d["edge_list"]["edges_per_step"])
except Exception as e:
    print("$v failed", e, open("$OUT/bench_c3_$v.err").read()[-1500:])
PY
done
FM_EDGE_FORM=stream timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"edge|es_" --launch-skip 15 -c 6 --csv --log-file $OUT/edge_stream_launches.csv \
    python bench.py --config c3 --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 3 > /dev/null 2>&1; grep -v "^==" $OUT/edge_stream_launches.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -3
