#!/bin/bash
# like ab_env.sh but prints the end-to-end figure too: tools/ab_env_e2e.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  echo "== $var=$v"
  env $var=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c '
import sys,json
d=json.loads(sys.stdin.read()); print("value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1),"e2e_ms",round(d["e2e"]["ms_per_step"],3))'
done
