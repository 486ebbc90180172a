#!/bin/bash
# BASELINE.json configs 3-5 commit shapes on one GPU: STARK-shaped traces and the rows x cols x rate_bits sweep.
# usage (on the GPU box): bash tools/sweep.sh > gpurun_out/sweep.jsonl
for shape in "18 1271 1" "18 2502 1" "18 64 3" "18 512 1" "20 128 2" "20 256 2" "22 64 3" "22 128 1" "24 64 1"; do
  set -- $shape
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --log-n $1 --cols $2 --rate-bits $3 2>&1 | grep -E '^\{|Error|error|failed' | head -2
done
