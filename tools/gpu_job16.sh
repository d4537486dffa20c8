#!/bin/bash
# gpurun call: window kernel with single-step handling of flagged words: parity, full-size parity, then c3 / c4
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py tests/test_reference_suite.py -m gpu -q 2>&1 | tail -6) > gpurun_out/pytest_16.log 2>&1
tail -n 2 gpurun_out/pytest_16.log
timeout 600 python bench.py --steps 6 --warmup 2 --configs c3,c4,c5 --no-e2e --no-cpu-baseline > gpurun_out/bench_16.json 2> gpurun_out/bench_16.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_16.json").read().strip().splitlines()[-1])
for k,r in d["configs"].items(): print(k,"value",round(r["value"],1),"ms",round(r["ms_per_step"],2),"probe",round(r["roofline"]["kernel_ms_per_launch"],2),"window",round(r["other_kernels_ms_per_step"]["kmer_ranges_rows_stats"],2),"keeping",r["result"]["keeping"])
PY
