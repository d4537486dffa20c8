#!/bin/bash
# gpurun call 3: CLI / feeder tests, ncu --set full of the kernels under study (scale 0.25), CLI text->text throughput
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_text_feeder.py tests/test_cli.py -m gpu -q 2>&1 | tail -60) > gpurun_out/pytest_cli.log 2>&1
tail -8 gpurun_out/pytest_cli.log
B="python bench.py --scale 0.25 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
N="ncu --set full --clock-control none"
FL_PHRED_MODE=1 timeout 600 $N --import-source on -k regex:k_phred_score -s 1 -c 1 -f -o gpurun_out/ncu_phred_score $B --configs c2 > gpurun_out/ncu1.log 2>&1
FL_PHRED_MODE=2 timeout 600 $N -k 'regex:k_phred_(sum|win)' -s 2 -c 2 -f -o gpurun_out/ncu_phred_sumwin $B --configs c2 > gpurun_out/ncu2.log 2>&1
FL_FILTER_KIND=2 timeout 600 $N --import-source on -k regex:k_probe_paint -s 1 -c 1 -f -o gpurun_out/ncu_probe_plain $B --configs c3 > gpurun_out/ncu3.log 2>&1
FL_FILTER_KIND=3 timeout 600 $N -k regex:k_probe_paint -s 1 -c 1 -f -o gpurun_out/ncu_probe_minimizer $B --configs c3 > gpurun_out/ncu4.log 2>&1
timeout 600 $N --import-source on -k regex:k_kmer_window -s 1 -c 1 -f -o gpurun_out/ncu_window $B --configs c3 > gpurun_out/ncu5.log 2>&1
timeout 600 $N -k regex:k_probe_paint -s 1 -c 1 -f -o gpurun_out/ncu_probe_c5 $B --configs c5 > gpurun_out/ncu6.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/ncu1.log gpurun_out/ncu3.log gpurun_out/ncu5.log
timeout 900 python tools/cli_e2e.py --small-gbp 0.5 --large-gbp 5 --kmer-gbp 0.1 > gpurun_out/cli_e2e.json 2> gpurun_out/cli_e2e.err; tail -3 gpurun_out/cli_e2e.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/cli_e2e.json").read())
for c in d["cases"]:
    print(c["case"], c["bases"], c.get("stdout_identical"), c.get("speedup_wall"))
    for k,v in c.items():
        if isinstance(v, dict) and "seconds" in v: print("   ", k, round(v["seconds"],2), "s", round(v.get("gbases_per_s",0),3), "Gb/s", v.get("phases"), v.get("rc"))
PY
