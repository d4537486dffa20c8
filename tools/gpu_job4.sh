#!/bin/bash
# gpurun call 4: kernels touched since call 3 (window kernel, feeder I/O), probe L2-policy variants, CLI throughput
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_text_feeder.py tests/test_cli.py tests/test_gpu_parity.py tests/test_golden.py "tests/test_gpu_fullsize.py::test_config3_kmer_full_size" -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_4.log 2>&1
tail -6 gpurun_out/pytest_4.log
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 6 --warmup 2 --configs $CFG --no-e2e --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    for k,r in d["configs"].items():
        print("$tag",k,"value",round(r["value"],1),"ms",round(r["ms_per_step"],2),"probe",round(r["roofline"]["kernel_ms_per_launch"],2),"window",round(r["other_kernels_ms_per_step"]["kmer_ranges_rows_stats"],2))
except Exception as e:
    print("$tag failed", e, open("gpurun_out/bench_$tag.err").read()[-400:])
PY
}
CFG=c3,c4
run plain_m2 FL_FILTER_KIND=2 FL_PROBE_MODE=2
CFG=c3
run plain_m4 FL_FILTER_KIND=2 FL_PROBE_MODE=4
run mini_m2 FL_FILTER_KIND=3 FL_PROBE_MODE=2
run mini_m4 FL_FILTER_KIND=3 FL_PROBE_MODE=4
run mini_m2_persist FL_FILTER_KIND=3 FL_PROBE_MODE=2 FL_L2_PERSIST=1
run mini_m4_persist FL_FILTER_KIND=3 FL_PROBE_MODE=4 FL_L2_PERSIST=1
run plain_m2_persist FL_FILTER_KIND=2 FL_PROBE_MODE=2 FL_L2_PERSIST=1
CFG=c5
run c5_m2 FL_PROBE_MODE=2
run c5_m4 FL_PROBE_MODE=4
timeout 900 python tools/cli_e2e.py --small-gbp 0.5 --large-gbp 10 --kmer-gbp 0.1 > gpurun_out/cli_e2e.json 2> gpurun_out/cli_e2e.err; tail -n 3 gpurun_out/cli_e2e.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/cli_e2e.json").read())
for c in d["cases"]:
    print(c["case"], c["bases"], c.get("stdout_identical"), c.get("speedup_wall"))
    for k,v in c.items():
        if isinstance(v, dict) and "seconds" in v: print("   ", k, round(v["seconds"],2), "s", round(v.get("gbases_per_s",0),3), "Gb/s", v.get("phases"), v.get("rc"))
PY
