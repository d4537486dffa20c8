#!/bin/bash
# gpurun call: the new reference-file tests alone, full tracebacks kept
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_text_feeder.py -m gpu -q -k "kmers_add_text or reference_files or gzip_input" > gpurun_out/pytest_refs.log 2>&1
grep -n "Error\|assert\|^E " gpurun_out/pytest_refs.log | head -40
tail -5 gpurun_out/pytest_refs.log
