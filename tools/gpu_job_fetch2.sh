#!/bin/bash
# random 32-byte sector loads against tables of growing size: where does the rate fall (L2 capacity, TLB reach)?
mkdir -p gpurun_out
: > gpurun_out/sector_fetch_sizes.jsonl
for lg in 18 20 21 22 23 24 25 26 27 28; do timeout 120 tools/sector_fetch_probe $lg quick >> gpurun_out/sector_fetch_sizes.jsonl 2>&1; done
cat gpurun_out/sector_fetch_sizes.jsonl
