#!/bin/bash
set -u
OUT=gpurun_out/r02_n; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,sm__cycles_elapsed.max,sm__cycles_active.avg,smsp__inst_executed.sum,smsp__average_warp_latency_issue_stalled_no_instruction.ratio,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_lg_throttle.ratio,smsp__average_warp_latency_issue_stalled_imc_miss.ratio,smsp__average_warp_latency_issue_stalled_branch_resolving.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,launch__waves_per_multiprocessor,sm__ctas_launched.sum,sm__maximum_warps_per_active_cycle_pct --cache-control none --clock-control none -k regex:formation_logic --launch-skip 60 -c 16 --csv --log-file $OUT/logic_steps.csv python bench.py --config form --steps 100 --warmup 5 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_n/logic_steps.csv")) if len(r) > 10]
hdr = rows[0]; by = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(hdr, r)); by.setdefault(d["ID"], {})[d["Metric Name"]] = d["Metric Value"]
keys = list(next(iter(by.values())).keys())
print(" | ".join(k.replace("smsp__average_warp_latency_issue_stalled_", "st_").replace(".ratio", "")[-22:] for k in keys))
for k, v in by.items():
    print(k, " | ".join(v.get(m, "?")[:10] for m in keys))
PY
