#!/bin/bash
# A/B on the GPU box: parity subset, then bench phase times per (poseidon variant, tree fuse) setting
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
run() {
  echo "== $*"
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['roofline']['phase_ms'].items()})
    else: print(l.rstrip()[-300:])
"
}
for s in "$@"; do run $s; done
