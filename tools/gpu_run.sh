mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload kmer > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err; cut -c1-200 gpurun_out/bench_c3.json
timeout 600 python bench.py --workload kmer --trim-split --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -2 gpurun_out/bench_c4.err; cut -c1-200 gpurun_out/bench_c4.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload kmer --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_c3.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_probe_paint -s 1 -c 1 -o gpurun_out/prof_c3_probe_anchor python bench.py --workload kmer --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1; tail -2 gpurun_out/ncu_c3.log
timeout 600 python bench.py --workload kmer --reads 1250000 --bases 12.5e9 --genome-bases 2000000000 --target-bases 3.75e9 --steps 5 --no-cpu-baseline > gpurun_out/bench_c5share.json 2> gpurun_out/bench_c5share.err; tail -3 gpurun_out/bench_c5share.err; cut -c1-200 gpurun_out/bench_c5share.json
