#!/bin/bash
# A/B run on the GPU box over one environment variable: tools/ab_env.sh VAR v1 v2 ...  (bench phase times per value)
var=$1; shift
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value",round(d["value"],1),"ms",round(d["ms_per_step"],3),{k:round(v,3) for k,v in d["roofline"]["phase_ms"].items()})
    else: print(l.rstrip())
'
for v in "$@"; do
  echo "== $var=$v"
  env $var=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
done
