#!/bin/bash
# gpurun call: after removing the dead load-flavour variants: parity + the k-mer configurations
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py tests/test_cli.py -m gpu -q 2>&1 | tail -6) > gpurun_out/pytest_18.log 2>&1
tail -n 2 gpurun_out/pytest_18.log
timeout 600 python bench.py --steps 6 --warmup 2 --configs c3,c5 --no-e2e --no-cpu-baseline > gpurun_out/bench_18.json 2> gpurun_out/bench_18.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_18.json").read().strip().splitlines()[-1])
for k,r in d["configs"].items(): print(k,"value",round(r["value"],1),"ms",round(r["ms_per_step"],2),"probe",round(r["roofline"]["kernel_ms_per_launch"],2),"window",round(r["other_kernels_ms_per_step"]["kmer_ranges_rows_stats"],2),"keeping",r["result"]["keeping"])
PY
