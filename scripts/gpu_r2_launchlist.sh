#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2l}
timeout 600 python scripts/time_overlap.py 2>&1 | tail -1 | tee $OUT/${TAG}_overlap.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-graph --no-extras > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list exit $?"; tail -2 $OUT/${TAG}_ncu_bench.log | cut -c1-300
python - <<PY
import csv
rows=[r for r in csv.reader(open('$OUT/${TAG}_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:60]), {})[r[mi]]=r[vi]
last={}
for (i,k),m in sorted(d.items()):
    last[k]=m
for k,m in last.items(): print(k, m)
PY
