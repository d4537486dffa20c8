#!/bin/bash
# gpurun call: compute-sanitizer memcheck over the golden cases and a slice of the parity tests (small inputs)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/memcheck_golden.log python -m pytest tests/test_golden.py -m gpu -q -x > gpurun_out/memcheck_golden.out 2>&1; echo "golden rc=$?"
tail -n 3 gpurun_out/memcheck_golden.out; tail -n 4 gpurun_out/memcheck_golden.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/memcheck_parity.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "prefilter or batching or kmer_assembly_random or phred_random or text" > gpurun_out/memcheck_parity.out 2>&1; echo "parity rc=$?"
tail -n 3 gpurun_out/memcheck_parity.out; tail -n 4 gpurun_out/memcheck_parity.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/memcheck_feeder.log python -m pytest tests/test_text_feeder.py -m gpu -q -x -k "push_text" > gpurun_out/memcheck_feeder.out 2>&1; echo "feeder rc=$?"
tail -n 3 gpurun_out/memcheck_feeder.out; tail -n 4 gpurun_out/memcheck_feeder.log
