#!/bin/bash
# A/B run on the GPU box: parity tests of the hashing path, then bench phase times per Poseidon variant.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "poseidon or hash_no_pad or merkle or commit_from_values or golden" 2>&1 | tail -3
for v in "$@"; do
  echo "== variant $v"
  VX_POSEIDON_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),d['roofline']['phase_ms'])
    else: print(l.rstrip())
"
done
