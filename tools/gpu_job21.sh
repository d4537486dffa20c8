#!/bin/bash
# gpurun call: which reference-file case differs from the reference binary
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_text_feeder.py -m gpu -q -k "reference_files" > gpurun_out/pytest_refcases.log 2>&1
grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/pytest_refcases.log | cut -c1-300 | head -40
