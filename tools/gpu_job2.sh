#!/bin/bash
# gpurun call 2: GPU tests, then k-mer configs with the new kernels B, A/B of the pre-filter flavours
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for kind in 3 2; do
  FL_FILTER_KIND=$kind timeout 900 python bench.py --steps 10 --warmup 3 --configs c3,c4 --no-cpu-baseline > gpurun_out/bench_k${kind}.json 2> gpurun_out/bench_k${kind}.err
  tail -3 gpurun_out/bench_k${kind}.err
done
timeout 900 python bench.py --steps 10 --warmup 3 --configs c5 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
for pm in 1 2; do
  FL_PHRED_MODE=$pm timeout 600 python bench.py --steps 20 --warmup 3 --configs c2 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_pm${pm}.json 2> gpurun_out/bench_c2_pm${pm}.err
  tail -2 gpurun_out/bench_c2_pm${pm}.err
  python -c "
import json;d=json.loads(open('gpurun_out/bench_c2_pm${pm}.json').read().strip().splitlines()[-1]);print('c2 phred_mode ${pm}', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['result'])"
done
python - <<'PY'
import json
for f in ("bench_k3","bench_k2","bench_c5"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        for k,r in d["configs"].items():
            if "error" in r: print(f,k,r); continue
            print(f,k,"value",round(r["value"],1),"ms",round(r["ms_per_step"],2),"probe",round(r["roofline"]["kernel_ms_per_launch"],2),"B",r["other_kernels_ms_per_step"],"e2e",r["e2e"].get("value") if r["e2e"] else None, r["e2e"].get("error") if r["e2e"] else None)
    except Exception as e:
        print(f,"parse failed",e)
PY
# launch list of one c4 step (per-kernel times)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 1 --warmup 1 --configs c4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_c4.csv")) if len(r)>5]
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if not hdr: continue
    d=dict(zip(hdr,r))
    try: v=float(d["Metric Value"].replace(",",""))
    except: continue
    u=d.get("Metric Unit","")
    ms = v/1e6 if u.startswith("n") else (v/1e3 if u.startswith("u") else v)
    k=d["Kernel Name"].split("(")[0][:60]
    agg[k][0]+=1; agg[k][1]+=ms
for k,(n,ms) in sorted(agg.items(), key=lambda x:-x[1][1])[:16]: print("%-62s n=%3d total %.3f ms"%(k,n,ms))
PY
