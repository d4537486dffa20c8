#!/bin/bash
# gpurun call: pair-keyed probe kernel compiled for 4 / 5 / 6 blocks per SM (latency bound: long scoreboard 11.9 per issue)
mkdir -p gpurun_out
: > gpurun_out/probe_variants5.jsonl
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 6 --warmup 2 --configs c3 --no-e2e --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    out={"variant":"$tag","env":"$*","configs":{}}
    for k,r in d["configs"].items():
        out["configs"][k]={"value":r["value"],"ms_per_step":r["ms_per_step"],"probe_ms":r["roofline"]["kernel_ms_per_launch"],"keeping":r["result"]["keeping"]}
        print("$tag",k,"value",round(r["value"],1),"ms",round(r["ms_per_step"],2),"probe",round(r["roofline"]["kernel_ms_per_launch"],2),"keeping",r["result"]["keeping"])
    open("gpurun_out/probe_variants5.jsonl","a").write(json.dumps(out)+"\n")
except Exception as e:
    print("$tag failed", e, open("gpurun_out/bench_$tag.err").read()[-400:])
PY
}
run occ4 FL_PROBE_OCC=4
run occ5 FL_PROBE_OCC=5
run occ6 FL_PROBE_OCC=6
run occ4b FL_PROBE_OCC=4
