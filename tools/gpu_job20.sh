#!/bin/bash
# gpurun call: wrapped-FASTA references on the device: the whole GPU suite (full log kept), smoke, reference-file timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1
grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/pytest_gpu_full.log | tail -15
grep -n "^E " gpurun_out/pytest_gpu_full.log | head -30
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time timeout 400 python tools/cli_gz.py --gbp 0 --short-reads 4000000 --assembly-gbp 1.0 --tmp /dev/shm > gpurun_out/cli_refs.json 2> gpurun_out/cli_refs.err) 2>&1 | grep real
tail -3 gpurun_out/cli_refs.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/cli_refs.json").read().strip().splitlines()[-1])
    for sect in ("reference_files", "assembly_file"):
        r = d.get(sect, {})
        print(sect, {k: v for k, v in r.items() if not isinstance(v, dict)})
        for k in ("device_text", "host_reader"):
            if k in r:
                print("  ", k, round(r[k]["seconds"], 2), "s rc", r[k]["rc"], [p for p in r[k]["phases"] if "reference" in p or "total" in p])
except Exception as e:
    print("parse failed", e)
PY
