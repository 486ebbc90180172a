#!/bin/bash
# quick A/B: GPU parity tests of the touched area + phase times of the bench workload
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['roofline']['phase_ms'].items()})"
