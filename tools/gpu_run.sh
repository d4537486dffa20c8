mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_c2_chk.json 2> gpurun_out/bench_c2_chk.err; tail -2 gpurun_out/bench_c2_chk.err
timeout 300 python bench.py --workload kmer --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_chk.json 2> gpurun_out/bench_c3_chk.err; tail -2 gpurun_out/bench_c3_chk.err
python - <<'PY'
import json
for f in ('bench_c2_chk','bench_c3_chk'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['traffic'], round(d['roofline']['frac'],3), d['e2e'] and round(d['e2e']['value'],1))
PY
timeout 600 python tools/kbuild_bench.py > gpurun_out/kbuild.jsonl 2> gpurun_out/kbuild.err; tail -3 gpurun_out/kbuild.err; cut -c1-700 gpurun_out/kbuild.jsonl
