"""Times MerkleTree::<F, PoseidonBN128Hash>::new on the device (wrap-circuit shapes): python tools/bn128_bench.py"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import vectorx_b200 as vx  # noqa: E402

ctx = vx.default_context(0)
rng = np.random.default_rng(1)
out = []
for log_n, w in [(15, 135), (16, 135), (15, 20), (12, 135)]:
    leaves = rng.integers(0, 0xFFFFFFFF00000001, size=(1 << log_n, w), dtype=np.uint64)
    ts = []
    for _ in range(4):
        t = time.perf_counter()
        tree = vx.MerkleTree.new(leaves, 4, ctx=ctx, hasher=vx.POSEIDON_BN128_HASH)
        ctx.sync()
        ts.append(time.perf_counter() - t)
        tree.close()
    perms = (1 << log_n) * (-(-w // 9)) + (1 << log_n) - 16
    out.append({"leaves": 1 << log_n, "width": w, "ms": min(ts) * 1e3, "permutations": perms,
                "Mperm_per_s": perms / min(ts) / 1e6})
    print(json.dumps(out[-1]), flush=True)
