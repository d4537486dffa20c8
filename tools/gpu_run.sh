mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err; cut -c1-300 gpurun_out/bench_c2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload phred --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_c2.out 2>&1
timeout 600 python bench.py --workload kmer > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err; cut -c1-300 gpurun_out/bench_c3.json
timeout 600 python bench.py --workload kmer --trim-split --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -2 gpurun_out/bench_c4.err; cut -c1-300 gpurun_out/bench_c4.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload kmer --trim-split --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_c4.out 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; tail -2 gpurun_out/bench_c2_ref.err; cut -c1-300 gpurun_out/bench_c2_ref.json
