#!/bin/bash
# A/B run on the GPU box: hybrid partial rounds (leaf-hash variant 8) -- parity of the hashing path per K, then bench phase
# times.  Usage: tools/ab_hybrid.sh K1 K2 ...
mkdir -p gpurun_out
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value",round(d["value"],1),"ms",round(d["ms_per_step"],3),{k:round(v,3) for k,v in d["roofline"]["phase_ms"].items()})
    else: print(l.rstrip())
'
echo "== default"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
for k in "$@"; do
  echo "== variant 8, $k spec-form partial rounds"
  VX_POSEIDON_VARIANT=8 VX_POSEIDON_NAIVE_ROUNDS=$k timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "commit_from_values or golden" 2>&1 | tail -1
  VX_POSEIDON_VARIANT=8 VX_POSEIDON_NAIVE_ROUNDS=$k timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
done
