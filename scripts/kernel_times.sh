#!/bin/bash
# per-kernel device times of bench steps (ncu, serialised launches): usage scripts/kernel_times.sh <out.csv> [rows-per-step] [level]
R=${2:-32768}; Z=${3:-2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --rows-per-step $R --level $Z"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:k_ -s 24 -c 24 --csv --log-file "$1" $CMD > gpurun_out/ncu1.log 2>&1
python - "$1" <<PY
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
gi=h.index("Grid Size") if "Grid Size" in h else None
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]),r[ki].split("(")[0],r[gi] if gi is not None else ""),{})[r[mi].split("__")[1][:22]]=r[vi]
for k,v in sorted(d.items()): print(k, v)
PY
