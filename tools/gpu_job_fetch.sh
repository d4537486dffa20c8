#!/bin/bash
# one random 32-byte sector load: DRAM bytes and time by flavour of the load (tools/sector_fetch_probe.cu)
mkdir -p gpurun_out
timeout 300 tools/sector_fetch_probe 26 > gpurun_out/sector_fetch.jsonl 2>&1
cat gpurun_out/sector_fetch.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_op_read.sum,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file gpurun_out/sector_fetch_ncu.csv tools/sector_fetch_probe 26 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/sector_fetch_ncu.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
cur={}
for r in rows[1:]:
    cur.setdefault(r[ii],{"k":r[ki][:40]})[r[mi]]=r[vi]
for i,d in cur.items(): print(i,d)
PY
