#!/bin/bash
# gpurun call: pair-keyed probe with all 16 filter words of a step requested up front
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q 2>&1 | tail -4) > gpurun_out/pytest_14.log 2>&1
tail -n 2 gpurun_out/pytest_14.log
for i in 1 2; do
timeout 600 python bench.py --steps 6 --warmup 2 --configs c3 --no-e2e --no-cpu-baseline > gpurun_out/bench_14_$i.json 2> gpurun_out/bench_14_$i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_14_$i.json").read().strip().splitlines()[-1])
for k,r in d["configs"].items(): print("run $i",k,"value",round(r["value"],1),"ms",round(r["ms_per_step"],2),"probe",round(r["roofline"]["kernel_ms_per_launch"],2),"keeping",r["result"]["keeping"])
PY
done
