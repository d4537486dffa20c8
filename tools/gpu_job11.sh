#!/bin/bash
# gpurun call: whole GPU suite + smoke, the default bench line, ncu --set full of the k-mer kernels (scale 0.25), CLI throughput
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time timeout 1200 python bench.py > gpurun_out/bench_all.json 2> gpurun_out/bench_all.err) 2>&1 | grep real
tail -3 gpurun_out/bench_all.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_all.json").read().strip().splitlines()[-1])
    print("c2", round(d["value"],1), d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d.get("cpu_baseline",{}).get("value"), d["gpu_launches"])
    for k,r in d["configs"].items():
        if "error" in r: print(k, r); continue
        print(k, round(r["value"],1), round(r["ms_per_step"],2), "probe", round(r["roofline"]["kernel_ms_per_launch"],2), "frac", round(r["roofline"]["frac"],3), "dram_frac", r["roofline"]["dram_frac"], r["other_kernels_ms_per_step"], "e2e", round(r["e2e"]["value"],1), "cpu", r.get("cpu_baseline",{}).get("value"), "build", r["kmers_build"]["build_ms_rank0"])
        print("   ", json.dumps(r["roofline"].get("request_rate_bounds"))[:500])
except Exception as e:
    print("bench parse failed", e)
PY
B="python bench.py --steps 1 --warmup 1 --scale 0.25 --no-e2e --no-cpu-baseline"
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:k_kmer_window -s 1 -c 1 -f -o gpurun_out/ncu_window_c4 $B --configs c4 > gpurun_out/ncu_w.log 2>&1
timeout 600 $N -k regex:k_kmer_scan -s 2 -c 2 -f -o gpurun_out/ncu_scan_c4 $B --configs c4 > gpurun_out/ncu_s.log 2>&1
timeout 600 $N -k regex:k_probe_paint -s 1 -c 1 -f -o gpurun_out/ncu_probe_c3_pair $B --configs c3 > gpurun_out/ncu_p.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 1500 python tools/cli_e2e.py --small-gbp 0.2 --large-gbp 20 --kmer-gbp 0.05 --tmp /dev/shm > gpurun_out/cli_e2e.json 2> gpurun_out/cli_e2e.err; tail -n 3 gpurun_out/cli_e2e.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/cli_e2e.json").read())
for c in d["cases"]:
    print(c["case"], c["bases"], c.get("stdout_identical"), c.get("speedup_wall"), c.get("generate_seconds"))
    for k,v in c.items():
        if isinstance(v, dict) and "seconds" in v: print("   ", k, round(v["seconds"],2), "s", round(v.get("gbases_per_s",0),3), "Gb/s", v.get("phases"), v.get("rc"))
PY
