mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --workload kmer > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; tail -3 gpurun_out/bench_c3_n2.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_c3_n2.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['result'],d['e2e'],d['roofline']['kernel_ms_per_launch'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --no-e2e > gpurun_out/bench_c2_n2b.json 2> gpurun_out/bench_c2_n2b.err; tail -3 gpurun_out/bench_c2_n2b.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_c2_n2b.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['result'])"
