#!/bin/bash
# gpurun call: final check of the round: whole GPU suite, smoke, the default bench line
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
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
