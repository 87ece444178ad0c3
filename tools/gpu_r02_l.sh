#!/bin/bash
set -u
OUT=gpurun_out/r02_l; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum --cache-control none --clock-control none -k regex:formation_logic --launch-skip 60 -c 30 --csv --log-file $OUT/logic_steps.csv python bench.py --config form --steps 100 --warmup 5 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_l/logic_steps.csv")) if len(r) > 10]
hdr = rows[0]; by = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(hdr, r)); by.setdefault(d["ID"], {})[d["Metric Name"]] = d["Metric Value"]
for k, v in by.items():
    print(k, " ".join(f"{m.split('__')[-1][:28]}={x}" for m, x in v.items()))
PY
