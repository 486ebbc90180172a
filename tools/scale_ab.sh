#!/bin/bash
# usage: tools/scale_ab.sh N   -- bench.py at N GPUs with the streamed sharded commit on and off
N=$1
for m in 1 0; do
  VX_SHARD_STREAM=$m timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/scale_ab.err | tail -1 > gpurun_out/bench_g${N}_stream${m}.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_g${N}_stream${m}.json").read())
    print("N=$N stream=$m", "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "e2e_ms", round(d["e2e"]["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["phase_ms"].items()})
except Exception as e:
    print("N=$N stream=$m failed:", e, open("gpurun_out/scale_ab.err").read()[-1500:])
PY
done
