#!/bin/bash
# usage: tools/scale.sh "8 4 2" [peer|nccl]   -- bench.py at the listed GPU counts, one summary line each
for N in $1; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --exchange ${2:-peer} 2>&1 | tail -1 > gpurun_out/bench_g${N}_${2:-peer}.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_g${N}_${2:-peer}.json").read())
    print("N=$N ${2:-peer}", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["phase_ms"].items()})
except Exception as e:
    print("N=$N failed:", open("gpurun_out/bench_g${N}_${2:-peer}.json").read()[-1500:])
PY
done
