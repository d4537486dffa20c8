#!/bin/bash
# gpurun call: the deferred second half of --trim / --split batches (fl_reads_push): parity, then config 4 with its e2e leg
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_cli.py tests/test_text_feeder.py tests/test_reference_suite.py -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_7.log 2>&1
tail -4 gpurun_out/pytest_7.log
timeout 900 python bench.py --steps 6 --warmup 3 --configs c3,c4 --no-cpu-baseline > gpurun_out/bench_7.json 2> gpurun_out/bench_7.err
tail -3 gpurun_out/bench_7.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_7.json").read().strip().splitlines()[-1])
for k,r in d["configs"].items():
    if "error" in r: print(k, r); continue
    print(k, round(r["value"],1), round(r["ms_per_step"],2), "probe", round(r["roofline"]["kernel_ms_per_launch"],2), r["other_kernels_ms_per_step"], "e2e", r["e2e"])
PY
