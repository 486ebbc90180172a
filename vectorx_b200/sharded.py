"""One commit sharded over G GPUs (coset partition) -- host-side plan and driver.

Rank g of G (G a power of two, G <= 2^cap_height) ends up owning leaf block [g N/G, (g+1) N/G): whole cap subtrees, so no
interior hashing crosses GPUs.  While G <= 2^rate_bits the block is a set of whole LDE cosets; beyond that (a rate_bits = 1
STARK trace on 8 GPUs) it is one of the 2^fold_bits equal parts of ONE coset's leaf block, which is itself the transform of
the coefficients folded onto a smaller coset (csrc/ntt.cu lde_fold_kernel) -- still no exchange beyond the coefficients.

    1. iNTT of this rank's column slice                       (no communication)
    2. all-gather of the coefficients, 8 n c bytes in total    (the ONE exchange step; NCCL over NVLink)
    3. coset NTTs + leaf hashing + subtrees of the own block   (no communication)
    4. all-gather of the 2^cap_height / G local cap entries    (512 B)

Two drivers share the plan: `sharded_commit` below keeps the exchange in a communicator object (torch.distributed:
NCCL on GPUs, gloo in the CPU tests), and `PeerGroup` hands the whole commit to the library, whose own kernels do the
exchange over NVLink peer memory (csrc/shard.cu) -- the path bench.py times at N > 1.

The compute steps go through an engine object: `DeviceEngine` binds the C ABI (libvectorx_b200.so) and is the only
engine in the product; tests inject a CPU engine to exercise the plan and the collective plumbing over gloo.
The path this replaces is plonky2's single-process PolynomialBatch::from_values (reached from
contracts/lib/succinctx/plonky2x/core/src/backend/circuit/build.rs:69-75); upstream has no multi-device form.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass


@dataclass(frozen=True)
class ShardPlan:
    world: int
    rank: int
    c: int
    log_n: int
    rate_bits: int
    cap_height: int

    def __post_init__(self):
        g = self.world
        if g < 1 or g & (g - 1):
            raise ValueError(f"world size {g} is not a power of two")
        if not 0 <= self.rank < g:
            raise ValueError(f"rank {self.rank} out of range for world size {g}")
        if g > (1 << self.cap_height) or g > (1 << (self.rate_bits + self.log_n)):
            raise ValueError(f"{g} shards need cap_height >= {g.bit_length() - 1} (whole cap subtrees per shard)")
        if self.c < 1:
            raise ValueError("empty batch")

    # ---- column slice for the iNTT stage (last ranks zero-padded so every rank gathers equal blocks)
    @property
    def cols_per_rank(self) -> int:
        return (self.c + self.world - 1) // self.world

    @property
    def col_lo(self) -> int:
        return min(self.c, self.rank * self.cols_per_rank)

    @property
    def col_hi(self) -> int:
        return min(self.c, self.col_lo + self.cols_per_rank)

    # ---- leaf block / cap slice owned after the exchange
    @property
    def n(self) -> int:
        return 1 << self.log_n

    @property
    def lde_size(self) -> int:
        return self.n << self.rate_bits

    @property
    def leaves(self) -> int:
        return self.lde_size // self.world

    @property
    def leaf_first(self) -> int:
        return self.rank * self.leaves

    @property
    def caps(self) -> int:
        return (1 << self.cap_height) // self.world

    @property
    def cap_first(self) -> int:
        return self.rank * self.caps

    @property
    def fold_bits(self) -> int:
        """log2 of the parts one coset's leaf block is split into (0 while every rank owns whole cosets)"""
        return max(0, (self.world.bit_length() - 1) - self.rate_bits)

    @property
    def fold_index(self) -> int:
        """which part of its coset's leaf block this rank owns"""
        return self.rank & ((1 << self.fold_bits) - 1)

    @property
    def cosets(self):
        """natural coset indices rho (LDE point i = 2^rate_bits k + rho) of the leaf blocks this rank owns (or owns a part
        of, when fold_bits > 0), in leaf order"""
        if self.fold_bits:
            b = self.rank >> self.fold_bits
            return [int(format(b, f"0{self.rate_bits}b")[::-1], 2) if self.rate_bits else 0]
        per = (1 << self.rate_bits) // self.world
        out = []
        for b in range(self.rank * per, (self.rank + 1) * per):
            rho = int(format(b, f"0{self.rate_bits}b")[::-1], 2) if self.rate_bits else 0
            out.append(rho)
        return out

    @property
    def exchange_bytes_per_rank(self) -> int:
        """bytes this rank receives in the coefficient all-gather"""
        return 8 * self.n * self.cols_per_rank * (self.world - 1)


class DeviceEngine:
    """Compute steps through the C ABI on this rank's GPU. Arrays are torch CUDA tensors (int64 views of u64)."""

    def __init__(self, ctx):
        from ._lib import check, load
        self.ctx, self.lib, self.check = ctx, load(), check

    def fence(self):
        """collectives run on torch's current stream, the library on its own: order them"""
        import torch
        torch.cuda.current_stream().synchronize()

    def intt(self, values, coeffs_out, ncols: int, log_n: int):
        self.check(self.lib.vx_ntt(self.ctx.handle, values.data_ptr(), coeffs_out.data_ptr(), ncols, log_n, 1, 0), "vx_ntt")

    def commit_shard(self, coeffs_all, plan: ShardPlan, cap_out):
        from ._lib import vp
        h = vp()
        self.check(self.lib.vx_commit_from_coeffs_shard(self.ctx.handle, coeffs_all.data_ptr(), plan.c, plan.log_n,
                                                        plan.rate_bits, plan.cap_height, plan.rank, plan.world,
                                                        ctypes.byref(h)), "vx_commit_from_coeffs_shard")
        self.check(self.lib.vx_batch_cap(h, cap_out.data_ptr()), "vx_batch_cap")
        return h

    def free(self, handle):
        self.lib.vx_batch_free(handle)


class TorchComm:
    """torch.distributed all-gather (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, dist):
        self.dist = dist

    def all_gather(self, out, inp):
        self.dist.all_gather_into_tensor(out, inp)


def sharded_commit(values_local, plan: ShardPlan, engine, comm, bufs):
    """values_local: this rank's (cols_per_rank, n) slice of the values (zero rows beyond col_hi).
    bufs: dict of preallocated tensors on the engine's device:
        coeff_mine (cols_per_rank, n), coeff_all (world * cols_per_rank, n), cap_loc (caps, 4), cap_all (2^cap_height, 4).
    Returns (shard handle, cap_all) -- cap_all is plonky2's merkle_tree.cap of the WHOLE commitment on every rank."""
    engine.intt(values_local, bufs["coeff_mine"], plan.cols_per_rank, plan.log_n)
    comm.all_gather(bufs["coeff_all"], bufs["coeff_mine"])
    engine.fence()
    handle = engine.commit_shard(bufs["coeff_all"], plan, bufs["cap_loc"])
    comm.all_gather(bufs["cap_all"], bufs["cap_loc"])
    engine.fence()
    return handle, bufs["cap_all"]


class PeerGroup:
    """This rank's end of a library-level shard group (vx_shard_group_*): the exchange runs inside the library's own
    kernels over peer memory.  `dist` (torch.distributed, initialised) is used once, to swap the 64-byte CUDA IPC
    handles of the ranks' gather buffers; pass dist=None with `local_peers` for ranks living in one process."""

    def __init__(self, ctx, plan: ShardPlan, dist=None):
        import numpy as np
        from ._lib import check, load, vp
        self.ctx, self.plan, self.lib, self.check = ctx, plan, load(), check
        self._h = vp()
        check(self.lib.vx_shard_group_create(ctx.handle, plan.rank, plan.world, plan.c, plan.log_n, plan.rate_bits,
                                             plan.cap_height, ctypes.byref(self._h)), "vx_shard_group_create")
        if plan.world > 1 and dist is not None:
            mine = np.zeros(64, dtype=np.uint8)
            check(self.lib.vx_shard_group_ipc_handle(self._h, mine.ctypes.data), "vx_shard_group_ipc_handle")
            gathered = [None] * plan.world
            dist.all_gather_object(gathered, mine.tobytes())
            table = np.frombuffer(b"".join(gathered), dtype=np.uint8).copy()
            check(self.lib.vx_shard_group_connect_ipc(self._h, table.ctypes.data), "vx_shard_group_connect_ipc")

    def set_layout(self, mode: int):
        """0 = by size (default), 1 = contiguous column slices, 2 = interleaved parts; same on every rank, before the
        first commit"""
        self.check(self.lib.vx_shard_group_set_layout(self._h, mode), "vx_shard_group_set_layout")

    def column_map(self):
        """global column index of each of this rank's cols_per_rank local columns (-1 = padding): the columns of the
        trace that make up `values_local`"""
        import numpy as np
        out = np.zeros(self.plan.cols_per_rank, dtype=np.uint32)
        self.check(self.lib.vx_shard_group_column_map(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))),
                   "vx_shard_group_column_map")
        return [(-1 if int(g) == 0xFFFFFFFF else int(g)) for g in out]

    def local_slice(self, full):
        """this rank's (cols_per_rank, n) slice of the full (c, n) value matrix, per column_map()"""
        import numpy as np
        mine = np.zeros((self.plan.cols_per_rank, full.shape[1]), dtype=np.uint64)
        for i, g in enumerate(self.column_map()):
            if g >= 0:
                mine[i] = full[g]
        return mine

    @staticmethod
    def connect_local(groups):
        """all ranks in this process: map each other's buffers directly (peer access)"""
        from ._lib import check, load, vp
        arr = (vp * len(groups))(*[g._h for g in groups])
        check(load().vx_shard_group_connect_local(arr, len(groups)), "vx_shard_group_connect_local")

    def commit_from_values(self, values_local, cap_all_out=None):
        """values_local: (cols_per_rank, n) slice, host (numpy / pinned torch) or device (torch).  Returns the shard
        handle (free with .free_batch); cap_all_out (2^cap_height, 4) receives the whole cap."""
        from ._lib import ptr, vp
        h = vp()
        self.check(self.lib.vx_shard_commit_from_values(self._h, ptr(values_local), ptr(cap_all_out), ctypes.byref(h)),
                   "vx_shard_commit_from_values")
        return h

    def free_batch(self, h):
        self.lib.vx_batch_free(h)

    def close(self):
        if self._h:
            self.lib.vx_shard_group_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
