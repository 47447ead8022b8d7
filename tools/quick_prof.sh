#!/bin/bash
# Instruction count and issue utilisation of the sampling kernel on a 500 k-document C4 slice
# (sweep 2 from random init): plain run first, then the same command under ncu (few counters).
#   tools/quick_prof.sh <tag>
tag=${1:-x}
CMD="python bench.py --docs 500000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0"
$CMD > gpurun_out/qp_${tag}_plain.log 2>&1 &&
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:k_gibbs_sweep -s 5 -c 5 --csv --log-file gpurun_out/qp_${tag}.csv $CMD > gpurun_out/qp_${tag}_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/qp_${tag}.csv")) if len(r)>10]
h=rows[0]; i_n=h.index("Metric Name"); i_v=h.index("Metric Value"); i_k=h.index("Kernel Name"); i_id=h.index("ID")
agg={}
for r in rows[1:]:
    agg.setdefault(r[i_id],{"k":r[i_k]})[r[i_n]]=float(r[i_v].replace(",",""))
ti=sum(v["smsp__inst_executed.sum"] for v in agg.values())
for k,v in agg.items(): print(v["k"][:60], {a:b for a,b in v.items() if a!="k"})
print("inst/token", ti/44969101)
PY
