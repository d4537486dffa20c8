#!/bin/bash
# gpurun --gpus 8: the default bench line on 8 ranks (config 5 in full: 100 Gbp / 10 M reads vs the 3 Gbp assembly)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -4 gpurun_out/bench_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
    print("c2", d["value"], d["ms_per_step"], d["config"]["sharding"], d["e2e"])
    for k,r in d["configs"].items():
        if "error" in r: print(k, r); continue
        print(k, r["value"], r["ms_per_step"], r["collectives_per_step"], r["kmers_build"]["broadcast_and_table_build_ms"], r["e2e"].get("value"), r["e2e"].get("h2d_gbs_this_rank"))
except Exception as e:
    print("parse failed", e)
PY
