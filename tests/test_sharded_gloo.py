"""World-size-2/4 CPU tests (gloo) of the sharded-commit plan and its collective plumbing (vectorx_b200/sharded.py).

The compute steps are injected from the CPU oracle, computing every shard DIRECTLY from the coset identity
(leaf block b = bit-reversed n-point transform of (g w_N^rho)^m c_m, rho = bitrev(b)) -- so the test also checks that
identity against a whole-batch commit.  The GPU engine is exercised by tests/test_gpu_parity.py::test_sharded_commit_*.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

P = 0xFFFFFFFF00000001
G_GEN = 14293326489335486720
W32 = 7277203076849721926


class OracleEngine:
    """CPU stand-in for DeviceEngine: numpy views of CPU torch tensors, arithmetic from oracle/."""

    def __init__(self):
        import oracle
        self.o = oracle
        self.shards = {}

    def fence(self):
        pass

    @staticmethod
    def _np(t):
        return t.numpy().view(np.uint64)

    def intt(self, values, coeffs_out, ncols, log_n):
        v, out = self._np(values), self._np(coeffs_out)
        for j in range(ncols):
            out[j] = self.o.fft(v[j], inverse=True)

    def commit_shard(self, coeffs_all, plan, cap_out):
        co = self._np(coeffs_all)[: plan.c]
        n, N = plan.n, plan.lde_size
        w_N = pow(W32, 1 << (32 - plan.log_n - plan.rate_bits), P)
        rows = np.zeros((plan.leaves, plan.c), dtype=np.uint64)
        bits = plan.log_n
        rev = np.array([int(format(i, f"0{bits}b")[::-1], 2) if bits else 0 for i in range(n)])
        part = n >> plan.fold_bits                            # fold_bits > 0: one part of ONE coset's leaf block
        for bi, rho in enumerate(plan.cosets):
            shift = G_GEN * pow(w_N, rho, P) % P
            for j in range(plan.c):
                ev = self.o.fft(co[j], shift=shift)          # evaluations on shift * <w_n>, natural order
                blk = ev[rev]                                 # leaf order inside the block
                if plan.fold_bits:
                    rows[:, j] = blk[plan.fold_index * part: (plan.fold_index + 1) * part]
                else:
                    rows[bi * n: (bi + 1) * n, j] = blk
        cap_loc_h = plan.cap_height - (plan.world.bit_length() - 1)
        digests, cap = self.o.merkle_new(rows, cap_loc_h)
        self._np(cap_out)[:] = cap
        self.shards[plan.rank] = (rows, digests)
        return plan.rank

    def free(self, handle):
        self.shards.pop(handle, None)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, c, log_n, rate, cap_h, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import oracle
        from vectorx_b200.sharded import ShardPlan, TorchComm, sharded_commit
        plan = ShardPlan(world, rank, c, log_n, rate, cap_h)
        n = 1 << log_n
        full = oracle.random_field((c, n), seed=4242)                    # same on every rank
        mine = np.zeros((plan.cols_per_rank, n), dtype=np.uint64)
        mine[: plan.col_hi - plan.col_lo] = full[plan.col_lo: plan.col_hi]
        t = lambda *shape: torch.zeros(shape, dtype=torch.int64)
        bufs = {"coeff_mine": t(plan.cols_per_rank, n), "coeff_all": t(world * plan.cols_per_rank, n),
                "cap_loc": t(plan.caps, 4), "cap_all": t(1 << cap_h, 4)}
        eng = OracleEngine()
        h, cap_all = sharded_commit(torch.from_numpy(mine.view(np.int64)), plan, eng, TorchComm(dist), bufs)
        whole = oracle.commit_from_values(full, rate, cap_h)
        ok_cap = np.array_equal(cap_all.numpy().view(np.uint64), whole["cap"])
        ok_coeffs = np.array_equal(bufs["coeff_all"].numpy().view(np.uint64)[:c], whole["coeffs"])
        rows, digests = eng.shards[h]
        ok_rows = np.array_equal(rows, whole["leaves"][plan.leaf_first: plan.leaf_first + plan.leaves])
        per_cap = whole["digests"].shape[0] // (1 << cap_h)
        ok_dig = np.array_equal(digests, whole["digests"][plan.cap_first * per_cap: (plan.cap_first + plan.caps) * per_cap])
        q.put((rank, ok_cap, ok_coeffs, ok_rows, ok_dig, plan.exchange_bytes_per_rank))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:                                                # surface the failure in the parent
        q.put((rank, "error", repr(e)))
        raise


@pytest.mark.parametrize("world,c,log_n,rate,cap_h", [(2, 9, 5, 3, 4), (2, 5, 4, 1, 1), (4, 7, 4, 2, 2),
                                                       (4, 6, 5, 1, 3)])      # more shards than cosets: folded blocks
def test_sharded_commit_over_gloo(world, c, log_n, rate, cap_h):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, c, log_n, rate, cap_h, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for res in results:
        assert res[1] != "error", res
        rank, ok_cap, ok_coeffs, ok_rows, ok_dig, xbytes = res
        assert ok_cap, f"rank {rank}: gathered cap differs from the whole-batch cap"
        assert ok_coeffs, f"rank {rank}: gathered coefficients differ"
        assert ok_rows, f"rank {rank}: shard leaves differ from the whole-batch slice"
        assert ok_dig, f"rank {rank}: shard digests differ from the whole-batch slice"
        assert xbytes == 8 * (1 << log_n) * ((c + world - 1) // world) * (world - 1)


def test_shard_plan_validation():
    from vectorx_b200.sharded import ShardPlan
    p = ShardPlan(8, 3, 135, 16, 3, 4)
    assert p.cols_per_rank == 17 and p.col_lo == 51 and p.col_hi == 68
    assert p.leaves == (1 << 19) // 8 and p.leaf_first == 3 * p.leaves and p.caps == 2 and p.cap_first == 6
    assert p.cosets == [6]                      # leaf block 3 = coset bitrev3(3) = 6
    assert ShardPlan(8, 7, 135, 16, 3, 4).col_hi == 135 and ShardPlan(8, 7, 135, 16, 3, 4).col_lo == 119
    assert ShardPlan(2, 1, 20, 10, 3, 4).cosets == [1, 5, 3, 7]
    q = ShardPlan(8, 5, 2502, 18, 1, 4)         # rate_bits = 1 STARK trace on 8 GPUs: quarter blocks of the two cosets
    assert q.fold_bits == 2 and q.fold_index == 1 and q.cosets == [1] and q.leaves == (1 << 19) // 8 and q.caps == 2
    assert ShardPlan(16, 0, 8, 4, 3, 4).fold_bits == 1
    for bad in [(3, 0, 8, 4, 3, 4), (32, 0, 8, 4, 3, 4), (4, 0, 8, 4, 3, 1), (2, 2, 8, 4, 3, 4), (2, 0, 0, 4, 3, 4)]:
        with pytest.raises(ValueError):
            ShardPlan(*bad)
