#!/bin/bash
# gpurun call: the reference-file tests with the group-parallel wrapped-FASTA gather, and the -a timing again
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_text_feeder.py -m gpu -q -k "reference_files or kmers_add_text or gzip_input" > gpurun_out/pytest_refcases2.log 2>&1
grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/pytest_refcases2.log | cut -c1-300 | head -30
timeout 80 python tools/cli_gz.py --gbp 0 --short-reads 200000 --assembly-gbp 1.0 --tmp /dev/shm > gpurun_out/cli_refs2.json 2> gpurun_out/cli_refs2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/cli_refs2.json").read().strip().splitlines()[-1])
    r = d.get("assembly_file", {})
    print({k: v for k, v in r.items() if not isinstance(v, dict)})
    for k in ("device_text", "host_reader"):
        if k in r:
            print("  ", k, round(r[k]["seconds"], 2), "s rc", r[k]["rc"], [p for p in r[k]["phases"] if "reference" in p or "total" in p])
except Exception as e:
    print("parse failed", e)
PY
