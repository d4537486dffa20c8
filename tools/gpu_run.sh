mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload phred > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err; cut -c1-400 gpurun_out/bench_c2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload phred --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_c2.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_phred_(sum|win)' -s 2 -c 2 -o gpurun_out/prof_c2_phred python bench.py --workload phred --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1; tail -2 gpurun_out/ncu_c2.log
timeout 600 python bench.py --workload kmer > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err; cut -c1-400 gpurun_out/bench_c3.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload kmer --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_c3.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_probe_paint -s 1 -c 1 -o gpurun_out/prof_c3_probe python bench.py --workload kmer --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1; tail -2 gpurun_out/ncu_c3.log
timeout 600 python bench.py --workload kmer --trim-split --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -2 gpurun_out/bench_c4.err; cut -c1-300 gpurun_out/bench_c4.json
timeout 600 python bench.py --impl reference --workload phred --steps 2 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; tail -2 gpurun_out/bench_c2_ref.err; cut -c1-300 gpurun_out/bench_c2_ref.json
