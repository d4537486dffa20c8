#!/bin/bash
# one gpurun call: GPU test suite, smoke, the full default bench line (all BASELINE configs)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2 > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_all.json 2> gpurun_out/bench_all.err) 2>&1 | grep real
tail -5 gpurun_out/bench_all.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_all.json").read().strip().splitlines()[-1])
    print("c2", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"], d.get("cpu_baseline",{}).get("value"))
    for k,r in d["configs"].items():
        if "error" in r: print(k, r); continue
        print(k, r["value"], r["ms_per_step"], r["roofline"]["kernel_ms_per_launch"], r["roofline"]["frac"], r["other_kernels_ms_per_step"], r["kmers_build"], r["e2e"], r.get("cpu_baseline",{}).get("value"))
except Exception as e:
    print("bench parse failed", e)
PY
