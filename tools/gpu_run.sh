mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 500 python tools/cli_e2e.py > gpurun_out/cli_e2e.json 2> gpurun_out/cli_e2e.err; tail -3 gpurun_out/cli_e2e.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/cli_e2e.json").read())
for c in d["cases"]:
    print(c["case"],c["bases"],c["stdout_identical"],c.get("speedup_wall"))
    for k in ("reference_cpu","ours_gpu"): print("  ",k,c[k]["seconds"],c[k].get("phases"))
PY
