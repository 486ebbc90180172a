"""Library-level sharded commit (csrc/shard.cu): the exchange is the library's own kernels over peer memory.

world = 1 runs on one GPU (the push/flag/wait kernels still run, against the rank's own buffer); world = 2 needs two
GPUs in this process and is skipped otherwise (bench.py --gpus N covers the one-process-per-GPU IPC path and checks
the gathered cap against a single-GPU commit).  Reference: plonky2 PolynomialBatch::from_values reached from
contracts/lib/succinctx/plonky2x/core/src/backend/circuit/build.rs:69-75.
"""
import threading

import numpy as np
import pytest

import oracle
import vectorx_b200 as vx
from vectorx_b200.sharded import PeerGroup, ShardPlan

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


def test_peer_group_world1_matches_oracle(ctx):
    c, log_n, rate, cap = 21, 10, 3, 2
    cols = oracle.random_field((c, 1 << log_n), seed=77)
    want = oracle.commit_from_values(cols, rate, cap)
    g = PeerGroup(ctx, ShardPlan(1, 0, c, log_n, rate, cap))
    cap_out = np.zeros((1 << cap, 4), dtype=np.uint64)
    for _ in range(3):                                   # epochs advance, buffers are reused
        h = g.commit_from_values(cols, cap_out)
        assert np.array_equal(cap_out, want["cap"])
        g.free_batch(h)
    g.close()


@pytest.mark.parametrize("world", [2, 4])
def test_peer_group_same_process(world):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs in one process")
    c, log_n, rate, cap = 37, 12, 3, 4
    cols = oracle.random_field((c, 1 << log_n), seed=5)
    want = oracle.commit_from_values(cols, rate, cap)
    ctxs = [vx.Context(d) for d in range(world)]
    plans = [ShardPlan(world, r, c, log_n, rate, cap) for r in range(world)]
    groups = [PeerGroup(ctxs[r], plans[r]) for r in range(world)]
    PeerGroup.connect_local(groups)
    caps = [np.zeros((1 << cap, 4), dtype=np.uint64) for _ in range(world)]
    errs = []

    def run(r):
        try:
            p = plans[r]
            mine = np.zeros((p.cols_per_rank, 1 << log_n), dtype=np.uint64)
            mine[: p.col_hi - p.col_lo] = cols[p.col_lo:p.col_hi]
            for _ in range(2):
                h = groups[r].commit_from_values(mine, caps[r])
                groups[r].free_batch(h)
        except Exception as e:                            # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for r in range(world):
        assert np.array_equal(caps[r], want["cap"])
    for g in groups:
        g.close()
    for cx in ctxs:
        cx.close()
