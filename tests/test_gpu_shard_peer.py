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
        b = vx.PolynomialBatch(ctx, h)                   # the rank's shard behind the handle (here: the whole commitment)
        leaves, digests = b.download()
        assert np.array_equal(b.polynomials, want["coeffs"])
        assert np.array_equal(leaves, want["leaves"]) and np.array_equal(digests, want["digests"])
        b.close()
    g.close()


@pytest.mark.parametrize("c,log_n,rate,cap", [(3, 8, 1, 1), (8, 9, 3, 4), (9, 9, 2, 0), (135, 10, 3, 4), (64, 12, 1, 4)])
def test_peer_group_world1_shapes(ctx, c, log_n, rate, cap):
    """The streamed form (sub-blocks of 8 / 16 / rest columns, sponge state carried between chunks) over the shapes that
    move its boundaries: narrow leaves (no hashing), exactly one rate group, a 1-column tail, the wires shape, STARK rate."""
    cols = oracle.random_field((c, 1 << log_n), seed=c + log_n)
    want = oracle.commit_from_values(cols, rate, cap)
    g = PeerGroup(ctx, ShardPlan(1, 0, c, log_n, rate, cap))
    cap_out = np.zeros((1 << cap, 4), dtype=np.uint64)
    h = g.commit_from_values(cols, cap_out)
    assert np.array_equal(cap_out, want["cap"])
    b = vx.PolynomialBatch(ctx, h)
    leaves, digests = b.download()
    assert np.array_equal(leaves, want["leaves"]) and np.array_equal(digests, want["digests"])
    b.close()
    g.close()


@pytest.mark.parametrize("world,layout,c,log_n,rate", [(2, 0, 37, 12, 3), (4, 0, 37, 12, 3),
                                                       (2, 2, 37, 12, 3),      # interleaved parts, whole cosets
                                                       (2, 2, 300, 13, 1),     # interleaved parts, many of them
                                                       (4, 2, 100, 13, 1),     # interleaved + folded leaf blocks (4 > 2^1)
                                                       (8, 2, 135, 13, 3), (8, 0, 70, 13, 1)])
def test_peer_group_same_process(world, layout, c, log_n, rate):
    """Ranks as threads of one process (peer access mapped directly).  layout 2 forces the interleaved column ownership
    that big slices use (the parts of all ranks alternate in sponge order), also together with leaf blocks that are
    parts of a coset (more ranks than cosets)."""
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs in one process")
    cap = 4
    cols = oracle.random_field((c, 1 << log_n), seed=5)
    want = oracle.commit_from_values(cols, rate, cap)
    ctxs = [vx.Context(d) for d in range(world)]
    plans = [ShardPlan(world, r, c, log_n, rate, cap) for r in range(world)]
    groups = [PeerGroup(ctxs[r], plans[r]) for r in range(world)]
    for g in groups:
        g.set_layout(layout)
    PeerGroup.connect_local(groups)
    owned = sorted(gc for g in groups for gc in g.column_map() if gc >= 0)
    assert owned == list(range(c))                        # every column has exactly one owner
    caps = [np.zeros((1 << cap, 4), dtype=np.uint64) for _ in range(world)]
    errs = []

    N = (1 << log_n) << rate
    shard_ok = [False] * world

    def run(r):
        try:
            mine = groups[r].local_slice(cols)
            for it in range(2):
                h = groups[r].commit_from_values(mine, caps[r])
                if it == 0:
                    groups[r].free_batch(h)
                    continue
                b = vx.PolynomialBatch(ctxs[r], h)      # this rank's leaf block and cap subtrees
                leaves, digests = b.download()
                lo, hi = r * N // world, (r + 1) * N // world
                per = want["digests"].shape[0] // world
                shard_ok[r] = (np.array_equal(leaves, want["leaves"][lo:hi])
                               and np.array_equal(digests, want["digests"][r * per:(r + 1) * per])
                               and np.array_equal(b.polynomials, want["coeffs"]))
                b.close()
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
    assert all(shard_ok), shard_ok
    for g in groups:
        g.close()
    for cx in ctxs:
        cx.close()
