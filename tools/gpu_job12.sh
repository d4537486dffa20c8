#!/bin/bash
# gpurun call: feeder reader threads / chunk size on a 20 Gbp file (fast generator: I/O experiment)
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 1500 python tools/cli_e2e.py --large-only --fast-gen --large-gbp 20 --tmp /dev/shm --large-env "r8c128:FL_READERS=8,FL_CHUNK_MB=128;r16c128:FL_READERS=16,FL_CHUNK_MB=128;r16c64:FL_READERS=16,FL_CHUNK_MB=64;r24c64:FL_READERS=24,FL_CHUNK_MB=64;r32c32:FL_READERS=32,FL_CHUNK_MB=32;r12c128:FL_READERS=12,FL_CHUNK_MB=128" > gpurun_out/cli_readers.json 2> gpurun_out/cli_readers.err; tail -n 3 gpurun_out/cli_readers.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/cli_readers.json").read())
for c in d["cases"]:
    print(c["case"], c["bases"], c.get("generate_seconds"))
    for k,v in c.items():
        if isinstance(v, dict) and "seconds" in v: print("   ", k, round(v["seconds"],2), "s", round(v.get("gbases_per_s",0),3), "Gb/s", [p for p in v.get("phases") if "pass 1" in p or "total" in p], v.get("rc"))
PY
