mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_golden.py -x -q -k "kmer or golden or reference_harness or bloom" 2>&1 | tail -8) > gpurun_out/pytest_kmer.log 2>&1
cat gpurun_out/pytest_kmer.log
for cfg in "1 1" "0 1" "1 0"; do set -- $cfg
FL_FILTER=$1 FL_ANCHOR=$2 timeout 300 python bench.py --workload kmer --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/bench_c3_f$1a$2.json 2> gpurun_out/bench_c3_f$1a$2.err; tail -2 gpurun_out/bench_c3_f$1a$2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c3_f$1a$2.json').read().strip().splitlines()[-1])
print('C3 filter',$1,'anchor',$2,'value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'probe_ms',round(d['roofline']['kernel_ms_per_launch'],1),'frac',round(d['roofline']['frac'],3),'build',d.get('kmers_build'))
PY
done
for a in 1 0; do
FL_ANCHOR=$a timeout 600 python bench.py --workload kmer --reads 1250000 --bases 12.5e9 --genome-bases 2000000000 --target-bases 3.75e9 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c5share_a$a.json 2> gpurun_out/bench_c5share_a$a.err; tail -3 gpurun_out/bench_c5share_a$a.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c5share_a$a.json').read().strip().splitlines()[-1])
print('C5share anchor',$a,'value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'probe_ms',round(d['roofline']['kernel_ms_per_launch'],1),'frac',round(d['roofline']['frac'],3),d.get('kmers_build'))
PY
done
