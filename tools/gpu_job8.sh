#!/bin/bash
# gpurun call: WIDE dense-set probe (8 sectors in flight per thread) and the 4-bit pre-filter, parity first
mkdir -p gpurun_out
(FL_FILTER=0 FL_PROBE_WIDE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -k "kmer or golden or assembly or trim or split" 2>&1 | tail -4) > gpurun_out/pytest_wide.log 2>&1
echo "wide: $(tail -n 1 gpurun_out/pytest_wide.log)"
(FL_FILTER_KIND=18 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -k "kmer or golden or assembly or trim or split" 2>&1 | tail -4) > gpurun_out/pytest_kind18.log 2>&1
echo "kind18: $(tail -n 1 gpurun_out/pytest_kind18.log)"
: > gpurun_out/probe_variants3.jsonl
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 6 --warmup 2 --configs $CFG --no-e2e --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    out={"variant":"$tag","env":"$*","configs":{}}
    for k,r in d["configs"].items():
        out["configs"][k]={"value":r["value"],"ms_per_step":r["ms_per_step"],"probe_ms":r["roofline"]["kernel_ms_per_launch"],"window_ms":r["other_kernels_ms_per_step"]["kmer_ranges_rows_stats"],"keeping":r["result"]["keeping"],"bounds":r["roofline"].get("request_rate_bounds")}
        print("$tag",k,"value",round(r["value"],1),"ms",round(r["ms_per_step"],2),"probe",round(r["roofline"]["kernel_ms_per_launch"],2),"keeping",r["result"]["keeping"])
    open("gpurun_out/probe_variants3.jsonl","a").write(json.dumps(out)+"\n")
except Exception as e:
    print("$tag failed", e, open("gpurun_out/bench_$tag.err").read()[-400:])
PY
}
CFG=c5
run c5_narrow FL_PROBE_WIDE=0
run c5_wide FL_PROBE_WIDE=1
CFG=c3
run c3_k2 FL_FILTER_KIND=2
run c3_k18 FL_FILTER_KIND=18
run c3_k18_l23 FL_FILTER_KIND=18 FL_FILTER_LOG2_WORDS=23
run c3_pair4 FL_FILTER_KIND=26
