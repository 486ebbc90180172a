"""One rank's share of a sharded commit, timed on ONE GPU: phase times of vx_commit_from_coeffs_shard (coset NTTs, leaf
hashing and cap subtrees of leaf block s of G) for the bench shape.  Usage: shard_phases.py [G] [reps]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vectorx_b200 as vx
from vectorx_b200._lib import DeviceArray

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = vx.default_context(0)
rng = np.random.default_rng(5)
coeffs = rng.integers(0, 0xFFFFFFFF00000001, size=(135, 1 << 16), dtype=np.uint64)
d = DeviceArray.from_host(ctx, coeffs)
acc = {}
wall = []
for i in range(reps + 3):
    t = time.perf_counter()
    b = vx.PolynomialBatch.from_coeffs_shard(d, 3, 4, i % G, G, ctx=ctx)
    dt = (time.perf_counter() - t) * 1e3
    ph = ctx.phase_ms()
    b.close()
    if i >= 3:
        wall.append(dt)
        for k, v in ph.items():
            acc[k] = acc.get(k, 0.0) + v / reps
print(json.dumps({"G": G, "wall_ms": round(sum(wall) / len(wall), 3), "phase_ms": {k: round(v, 3) for k, v in acc.items()},
                  "env": {k: v for k, v in os.environ.items() if k.startswith("VX_")}}))
