mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "phred" 2>&1 | tail -8) > gpurun_out/pytest_phred.log 2>&1
cat gpurun_out/pytest_phred.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_x.json 2> gpurun_out/bench_c2_x.err; tail -2 gpurun_out/bench_c2_x.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c2_x.json').read().strip().splitlines()[-1])
print('C2 value',d['value'],'ms',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms_per_launch'],'frac',d['roofline']['frac'])
PY
for eb in 4 8 16; do
timeout 300 python bench.py --workload kmer --no-cpu-baseline --steps 3 --e2e-batches $eb > gpurun_out/bench_c3_eb$eb.json 2> gpurun_out/bench_c3_eb$eb.err; tail -2 gpurun_out/bench_c3_eb$eb.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c3_eb$eb.json').read().strip().splitlines()[-1])
print('C3 e2e batches',$eb,d['e2e'])
PY
done
