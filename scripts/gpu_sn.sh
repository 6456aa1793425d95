#!/bin/bash
cd "$(dirname "$0")/.."
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
for b in 32 128; do
timeout 600 ncu --metrics $M --clock-control none -k regex:"mlp_tc_kernel|vlad_tc_kernel|sa_narrow" -s 16 -c 8 --csv --log-file gpurun_out/b${b}_tc.csv python bench.py --batch $b --steps 1 --warmup 3 --mode eager --no-cpu-baseline --no-extras --repeats 1 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open("gpurun_out/b${b}_tc.csv")))
h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
col={n:i for i,n in enumerate(rows[h])}
d=collections.OrderedDict()
for r in rows[h+1:]:
    if r and r[0].isdigit():
        d.setdefault(r[0],{"k":r[col["Kernel Name"]][:28],"g":r[col["Grid Size"]]})[r[col["Metric Name"]]]=r[col["Metric Value"]]+" "+r[col["Metric Unit"]]
print("batch ${b}")
for k,v in d.items(): print(v["k"],v["g"],v.get("gpu__time_duration.sum"),"dramR",v.get("dram__bytes_read.sum"),"dramW",v.get("dram__bytes_write.sum"),"L2",v.get("lts__t_bytes.sum"),"hit",v.get("lts__t_sector_hit_rate.pct"),"pipe",v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"))
PY
done
