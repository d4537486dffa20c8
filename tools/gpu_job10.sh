#!/bin/bash
# gpurun call: parity after the window kernel's cached epoch limit, then launch lists of c4 and c5 (kernel shares + DRAM bytes)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_10.log 2>&1
tail -3 gpurun_out/pytest_10.log
for c in c4 c5; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$c.csv python bench.py --steps 2 --warmup 1 --configs $c --no-e2e --no-cpu-baseline > gpurun_out/ncu_l_$c.log 2>&1
done
python - <<'PY'
import csv,collections
for c in ("c4","c5"):
    rows=[r for r in csv.reader(open("gpurun_out/launches_%s.csv"%c)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
    t=collections.OrderedDict()
    for r in rows[1:]:
        if r[mi]=="gpu__time_duration.sum": t[int(r[ii])]=(r[ki][:50], float(r[vi].replace(',',''))/1e6)
    ids=sorted(t); last=[i for i in ids if 'k_tiles_of' in t[i][0]][-1]
    print(c, [(t[i][0][:28], round(t[i][1],2)) for i in ids if i>=last and t[i][1]>0.3])
PY
