#!/bin/bash
# gpurun call: gzip input + reference files through the device-text path: the whole GPU suite, then tools/cli_gz.py
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time timeout 600 python tools/cli_gz.py --tmp /dev/shm > gpurun_out/cli_gz.json 2> gpurun_out/cli_gz.err) 2>&1 | grep real
tail -3 gpurun_out/cli_gz.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/cli_gz.json").read().strip().splitlines()[-1])
    print("bases", d["bases"], "plain", d["plain_bytes"], "gz", d["gz_bytes"], "bgzf", d["bgzf_bytes"], "compress_s", round(d["compress_seconds"], 1), "cpus", d["host_cpus"])
    for k, v in d.items():
        if isinstance(v, dict) and "seconds" in v:
            print(k, round(v["seconds"], 2), "s", round(v["gbases_per_s"], 3), "Gb/s rc", v["rc"], v["phases"][:8])
    print("identical", d["all_outputs_identical"])
    r = d.get("reference_files", {})
    for k in ("device_text", "host_reader"):
        if k in r:
            print("ref", k, round(r[k]["seconds"], 2), r[k]["rc"], r[k]["phases"][:6])
    print("ref identical", r.get("outputs_identical"))
except Exception as e:
    print("parse failed", e)
PY
