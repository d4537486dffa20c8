#!/bin/bash
# gpurun call: group-keyed / pair-keyed pre-filter and the L2::64B table load against the shipped probe kernel
mkdir -p gpurun_out
for kind in 6 10; do
  (FL_FILTER_KIND=$kind timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -k "kmer or golden or assembly or trim or split" 2>&1 | tail -4) > gpurun_out/pytest_kind$kind.log 2>&1
  echo "kind $kind: $(tail -n 1 gpurun_out/pytest_kind$kind.log)"
done
: > gpurun_out/probe_variants2.jsonl
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 6 --warmup 2 --configs $CFG --no-e2e --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    out={"variant":"$tag","env":"$*","configs":{}}
    for k,r in d["configs"].items():
        out["configs"][k]={"value":r["value"],"ms_per_step":r["ms_per_step"],"probe_ms":r["roofline"]["kernel_ms_per_launch"],"window_ms":r["other_kernels_ms_per_step"]["kmer_ranges_rows_stats"],"keeping":r["result"]["keeping"]}
        print("$tag",k,"value",round(r["value"],1),"ms",round(r["ms_per_step"],2),"probe",round(r["roofline"]["kernel_ms_per_launch"],2),"keeping",r["result"]["keeping"])
    open("gpurun_out/probe_variants2.jsonl","a").write(json.dumps(out)+"\n")
except Exception as e:
    print("$tag failed", e, open("gpurun_out/bench_$tag.err").read()[-400:])
PY
}
CFG=c3
run k2_m2 FL_FILTER_KIND=2 FL_PROBE_MODE=2
run k2_m4 FL_FILTER_KIND=2 FL_PROBE_MODE=4
run g4_22_m2 FL_FILTER_KIND=6 FL_FILTER_LOG2_WORDS=22 FL_PROBE_MODE=2
run g4_23_m2 FL_FILTER_KIND=6 FL_FILTER_LOG2_WORDS=23 FL_PROBE_MODE=2
run g4_23_m4 FL_FILTER_KIND=6 FL_FILTER_LOG2_WORDS=23 FL_PROBE_MODE=4
run g4_24_m4 FL_FILTER_KIND=6 FL_FILTER_LOG2_WORDS=24 FL_PROBE_MODE=4
run pair_22_m2 FL_FILTER_KIND=10 FL_FILTER_LOG2_WORDS=22 FL_PROBE_MODE=2
run pair_22_m4 FL_FILTER_KIND=10 FL_FILTER_LOG2_WORDS=22 FL_PROBE_MODE=4
run pair_23_m4 FL_FILTER_KIND=10 FL_FILTER_LOG2_WORDS=23 FL_PROBE_MODE=4
run nofilter_m4 FL_FILTER=0 FL_PROBE_MODE=4
CFG=c5
run c5_m2 FL_PROBE_MODE=2
run c5_m4 FL_PROBE_MODE=4
