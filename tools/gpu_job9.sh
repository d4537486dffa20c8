#!/bin/bash
# gpurun call: ballot-counted k_kmer_scan + automatic pre-filter flavour: parity, then which flavour wins at 5 M / 10 M / 20 M members
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_cli.py tests/test_reference_suite.py -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_9.log 2>&1
tail -3 gpurun_out/pytest_9.log
for kind in 22 26; do
  (FL_FILTER_KIND=$kind timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -k "kmer or golden or assembly or trim or split" 2>&1 | tail -4) > gpurun_out/pytest_kind$kind.log 2>&1
  echo "kind $kind: $(tail -n 1 gpurun_out/pytest_kind$kind.log)"
done
: > gpurun_out/probe_variants4.jsonl
run() { # tag, extra bench args, env...
  tag=$1; shift; extra=$1; shift
  env "$@" timeout 600 python bench.py --steps 6 --warmup 2 --configs $CFG --no-e2e --no-cpu-baseline $extra > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    out={"variant":"$tag","env":"$*","args":"$extra","configs":{}}
    for k,r in d["configs"].items():
        out["configs"][k]={"value":r["value"],"ms_per_step":r["ms_per_step"],"probe_ms":r["roofline"]["kernel_ms_per_launch"],"window_ms":r["other_kernels_ms_per_step"]["kmer_ranges_rows_stats"],"keeping":r["result"]["keeping"],"n_kmers":r["kmers_build"]["n_kmers"]}
        print("$tag",k,"n_kmers",r["kmers_build"]["n_kmers"],"value",round(r["value"],1),"ms",round(r["ms_per_step"],2),"probe",round(r["roofline"]["kernel_ms_per_launch"],2),"window",round(r["other_kernels_ms_per_step"]["kmer_ranges_rows_stats"],2),"keeping",r["result"]["keeping"])
    open("gpurun_out/probe_variants4.jsonl","a").write(json.dumps(out)+"\n")
except Exception as e:
    print("$tag failed", e, open("gpurun_out/bench_$tag.err").read()[-400:])
PY
}
CFG=c3,c4
run auto_20m "" FL_X=0
CFG=c3
run k18_20m "" FL_FILTER_KIND=18
S10="--genome-bases 5e6 --illumina-pairs 2.5e6"
run k18_10m "$S10" FL_FILTER_KIND=18
run k26_10m "$S10" FL_FILTER_KIND=26
run k22_10m "$S10" FL_FILTER_KIND=22
S5="--genome-bases 2.5e6 --illumina-pairs 1.25e6"
run k18_5m "$S5" FL_FILTER_KIND=18
run k26_5m "$S5" FL_FILTER_KIND=26
run k22_5m "$S5" FL_FILTER_KIND=22
S15="--genome-bases 7.5e6 --illumina-pairs 3.75e6"
run k26_15m "$S15" FL_FILTER_KIND=26
run k22_15m "$S15" FL_FILTER_KIND=22
