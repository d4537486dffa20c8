#!/bin/bash
# gpurun --gpus 2: the NCCL path behind the C ABI (tests), the sharded CLI, the full bench on 2 ranks
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.txt 2>&1
(timeout 1200 python -m pytest tests/test_nccl_ranks.py "tests/test_text_feeder.py::test_cli_sharded_over_gpus_prints_what_one_gpu_prints" tests/test_gpu_parity.py tests/test_golden.py -m gpu -q 2>&1 | tail -40) > gpurun_out/pytest_n2.log 2>&1
tail -6 gpurun_out/pytest_n2.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -4 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
    print("c2", d["value"], d["ms_per_step"], d["config"]["sharding"], d["e2e"])
    for k,r in d["configs"].items():
        if "error" in r: print(k, r); continue
        print(k, r["value"], r["ms_per_step"], r["collectives_per_step"], r["kmers_build"]["broadcast_and_table_build_ms"], r["e2e"].get("value"))
except Exception as e:
    print("parse failed", e)
PY
