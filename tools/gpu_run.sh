mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "phred" 2>&1 | tail -25) > gpurun_out/pytest_phred.log 2>&1
cat gpurun_out/pytest_phred.log
for occ in 4 6; do
FL_PHRED_OCC=$occ timeout 300 python bench.py --workload phred --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/bench_phred_occ$occ.json 2> gpurun_out/bench_phred_occ$occ.err; tail -2 gpurun_out/bench_phred_occ$occ.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_phred_occ$occ.json').read().strip().splitlines()[-1])
print('occ',$occ,'value',d['value'],'ms',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms_per_launch'],'frac',d['roofline']['frac'])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_phred_(sum|win)' -c 2 -o gpurun_out/prof_phred_split python bench.py --workload phred --scale 0.1 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_tile.log 2>&1; tail -2 gpurun_out/ncu_tile.log
