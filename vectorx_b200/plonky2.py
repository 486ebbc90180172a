"""Host-side mirror of the plonky2 v0.2.0 objects VectorX's prover reaches on this path.

Same names and argument meaning as upstream (`PolynomialBatch::from_values`, `from_coeffs`,
`get_lde_values`, `MerkleTree::new`, `.cap`, `.prove`, `verify_merkle_proof_to_cap` ...), so the
parity tests read like the reference's own (e.g. the MerkleTree test at
contracts/lib/succinctx/plonky2x/core/src/backend/wrapper/poseidon_bn128.rs:209-267).  Every method
calls the C ABI; data stays on the device behind the handle and is materialised lazily.
The Rust binding a maintainer would write instead of this file is in INTEGRATION.md.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import Context, VxError, check, default_context, load, ptr, vp

P = 0xFFFFFFFF00000001

# Hasher of a GenericConfig: PoseidonGoldilocksConfig::Hasher = PoseidonHash; the wrap circuit's
# PoseidonBN128GoldilocksConfig::Hasher = PoseidonBN128Hash (P2X/backend/wrapper/plonky2_config.rs:19-27)
POSEIDON_HASH = 0
POSEIDON_BN128_HASH = 1


def reverse_bits(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


class MerkleCap:
    def __init__(self, hashes: np.ndarray):
        self.hashes = hashes              # (2^cap_height, 4) canonical u64

    def height(self) -> int:
        return int(self.hashes.shape[0]).bit_length() - 1

    def flatten(self) -> np.ndarray:
        return self.hashes.reshape(-1)

    def __eq__(self, other):
        return isinstance(other, MerkleCap) and np.array_equal(self.hashes, other.hashes)


class MerkleProof:
    def __init__(self, siblings: np.ndarray):
        self.siblings = siblings          # (depth, 4), bottom-up


class MerkleTree:
    """plonky2 hash/merkle_tree.rs MerkleTree<F, H>, H = PoseidonHash (default) or PoseidonBN128Hash."""

    def __init__(self, ctx: Context, handle, n: int, w: int, cap_height: int, hasher: int = POSEIDON_HASH):
        self._ctx, self._h, self.n, self.w, self.cap_height = ctx, handle, n, w, cap_height
        self.hasher = hasher
        self._cap = None

    @classmethod
    def new(cls, leaves, cap_height: int, ctx: Context | None = None, hasher: int = POSEIDON_HASH) -> "MerkleTree":
        """MerkleTree::<F, H>::new(leaves, cap_height); leaves is an (n, w) uint64 array (host) or tensor."""
        ctx = ctx or default_context()
        n, w = int(leaves.shape[0]), int(leaves.shape[1])
        if isinstance(leaves, np.ndarray):
            leaves = np.ascontiguousarray(leaves, dtype=np.uint64)
        h = vp()
        check(load().vx_merkle_new_hasher(ctx.handle, hasher, ptr(leaves), n, w, cap_height, None, None,
                                          ctypes.byref(h)), "vx_merkle_new")
        return cls(ctx, h, n, w, cap_height, hasher)

    @property
    def cap(self) -> MerkleCap:
        if self._cap is None:
            out = np.zeros((1 << self.cap_height, 4), dtype=np.uint64)
            check(load().vx_tree_cap(self._h, ptr(out)), "vx_tree_cap")
            self._cap = MerkleCap(out)
        return self._cap

    def get(self, i: int) -> np.ndarray:
        return self.get_many([i])[0]

    def get_many(self, idx) -> np.ndarray:
        idx = np.ascontiguousarray(np.array(idx, dtype=np.uint64))
        out = np.zeros((idx.size, self.w), dtype=np.uint64)
        check(load().vx_tree_leaves(self._h, ptr(idx), idx.size, ptr(out)), "vx_tree_leaves")
        return out

    def prove(self, leaf_index: int) -> MerkleProof:
        return self.prove_many([leaf_index])[0]

    def prove_many(self, idx) -> list:
        idx = np.ascontiguousarray(np.array(idx, dtype=np.uint64))
        depth = (self.n.bit_length() - 1) - self.cap_height
        out = np.zeros((idx.size, depth, 4), dtype=np.uint64)
        check(load().vx_tree_prove(self._h, ptr(idx), idx.size, ptr(out)), "vx_tree_prove")
        return [MerkleProof(out[i]) for i in range(idx.size)]

    def close(self):
        if self._h:
            load().vx_tree_free(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merkle_tree_digests(leaves: np.ndarray, cap_height: int, ctx: Context | None = None, hasher: int = POSEIDON_HASH):
    """MerkleTree::new returning the host-layout (digests, cap) -- what `.circuit` serialisation stores."""
    ctx = ctx or default_context()
    leaves = np.ascontiguousarray(leaves, dtype=np.uint64)
    n, w = leaves.shape
    digests = np.zeros((max(2 * (n - (1 << cap_height)), 0), 4), dtype=np.uint64)
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    check(load().vx_merkle_new_hasher(ctx.handle, hasher, ptr(leaves), n, w, cap_height,
                                      ptr(digests) if digests.size else None, ptr(cap), None), "vx_merkle_new")
    return digests, cap


class PolynomialBatch:
    """plonky2 fri/oracle.rs PolynomialBatch<F, C, D> (blinding = false, as every VectorX caller)."""

    def __init__(self, ctx: Context, handle):
        self._ctx, self._h = ctx, handle
        shape = (ctypes.c_uint32 * 4)()
        check(load().vx_batch_shape(handle, shape), "vx_batch_shape")
        self.num_polys, self.degree_log, self.rate_bits, self.cap_height = (int(x) for x in shape)
        self.blinding = False
        self._cap = None
        sh = (ctypes.c_uint64 * 3)()
        check(load().vx_batch_shard(handle, sh), "vx_batch_shard")
        self.leaf_first, self.num_leaves, self.num_caps = (int(x) for x in sh)   # == (0, N, 2^cap) unless sharded

    @classmethod
    def from_values(cls, values, rate_bits: int, blinding: bool, cap_height: int, timing=None,
                    fft_root_table=None, ctx: Context | None = None, hasher: int = POSEIDON_HASH) -> "PolynomialBatch":
        """values: (c, n) uint64 array/tensor, polynomial j = values[j] (evaluations on the subgroup) -- or, as plonky2
        holds them (`Vec<PolynomialValues<F>>`), a list of c separately allocated length-n arrays / tensors, which go to
        the device column by column without being flattened on the host."""
        return cls._commit("vx_commit_from_values_hasher", values, rate_bits, blinding, cap_height, ctx, hasher)

    @classmethod
    def from_coeffs(cls, polynomials, rate_bits: int, blinding: bool, cap_height: int, timing=None,
                    fft_root_table=None, ctx: Context | None = None, hasher: int = POSEIDON_HASH) -> "PolynomialBatch":
        return cls._commit("vx_commit_from_coeffs_hasher", polynomials, rate_bits, blinding, cap_height, ctx, hasher)

    @classmethod
    def from_values_keep(cls, values, values_dev, rate_bits: int, cap_height: int, ctx: Context | None = None) -> "PolynomialBatch":
        """from_values that also fills `values_dev` (a DeviceArray of the same shape) with the values: the wires commit of
        the prover, which reads the witness again for Z / partial products."""
        ctx = ctx or default_context()
        c, n = int(values.shape[0]), int(values.shape[1])
        log_n = n.bit_length() - 1
        if (1 << log_n) != n:
            raise VxError(f"polynomial length {n} is not a power of two")
        if isinstance(values, np.ndarray):
            values = np.ascontiguousarray(values, dtype=np.uint64)
        h = vp()
        check(load().vx_commit_from_values_keep(ctx.handle, ptr(values), c, log_n, rate_bits, cap_height, ptr(values_dev),
                                                ctypes.byref(h)), "vx_commit_from_values_keep")
        return cls(ctx, h)

    @classmethod
    def from_coeffs_shard(cls, polynomials, rate_bits: int, cap_height: int, shard_index: int, shard_count: int,
                          ctx: Context | None = None) -> "PolynomialBatch":
        """One GPU's share of a commit: leaves [s*N/S, (s+1)*N/S) from all coefficient columns."""
        ctx = ctx or default_context()
        c, n = int(polynomials.shape[0]), int(polynomials.shape[1])
        if isinstance(polynomials, np.ndarray):
            polynomials = np.ascontiguousarray(polynomials, dtype=np.uint64)
        h = vp()
        check(load().vx_commit_from_coeffs_shard(ctx.handle, ptr(polynomials), c, n.bit_length() - 1, rate_bits,
                                                 cap_height, shard_index, shard_count, ctypes.byref(h)),
              "vx_commit_from_coeffs_shard")
        return cls(ctx, h)

    @classmethod
    def _commit(cls, fn, data, rate_bits, blinding, cap_height, ctx, hasher=POSEIDON_HASH):
        if blinding:
            raise VxError("blinding=true is unsupported (zero_knowledge=false in standard_recursion_config)")
        ctx = ctx or default_context()
        if isinstance(data, (list, tuple)):
            # plonky2's own shape: Vec<PolynomialValues<F>> / Vec<PolynomialCoeffs<F>>, one allocation per column
            if hasher != POSEIDON_HASH:
                raise VxError("per-column input is bound for the Poseidon hasher only")
            cols = [np.ascontiguousarray(x, dtype=np.uint64) if isinstance(x, np.ndarray) else x for x in data]
            c = len(cols)
            n = int(cols[0].shape[0]) if c else 0
            log_n = n.bit_length() - 1
            if c == 0 or (1 << log_n) != n or any(int(x.shape[0]) != n or len(x.shape) != 1 for x in cols):
                raise VxError("columns must be non-empty 1-D arrays of one power-of-two length")
            ptrs = (vp * c)(*[ptr(x) for x in cols])
            name = "vx_commit_from_values_cols" if "values" in fn else "vx_commit_from_coeffs_cols"
            h = vp()
            check(getattr(load(), name)(ctx.handle, ptrs, c, log_n, rate_bits, cap_height, ctypes.byref(h)), name)
            return cls(ctx, h)
        c, n = int(data.shape[0]), int(data.shape[1])
        log_n = n.bit_length() - 1
        if (1 << log_n) != n:
            raise VxError(f"polynomial length {n} is not a power of two")
        if isinstance(data, np.ndarray):
            data = np.ascontiguousarray(data, dtype=np.uint64)
        h = vp()
        check(getattr(load(), fn)(ctx.handle, hasher, ptr(data), c, log_n, rate_bits, cap_height, ctypes.byref(h)), fn)
        return cls(ctx, h)

    # -- merkle_tree view -------------------------------------------------------------------------
    @property
    def n(self) -> int:
        return 1 << self.degree_log

    @property
    def lde_size(self) -> int:
        return 1 << (self.degree_log + self.rate_bits)

    @property
    def cap(self) -> MerkleCap:
        if self._cap is None:
            out = np.zeros((self.num_caps, 4), dtype=np.uint64)
            check(load().vx_batch_cap(self._h, ptr(out)), "vx_batch_cap")
            self._cap = MerkleCap(out)
        return self._cap

    @property
    def polynomials(self) -> np.ndarray:
        """coefficients, (c, n)."""
        out = np.zeros((self.num_polys, self.n), dtype=np.uint64)
        check(load().vx_batch_coeffs(self._h, ptr(out)), "vx_batch_coeffs")
        return out

    def get_lde_values(self, index: int, step: int = 1) -> np.ndarray:
        """leaves[reverse_bits(index * step, lde_bits)] (one LDE row, all polynomials)."""
        bits = self.degree_log + self.rate_bits
        return self.leaves([reverse_bits(index * step, bits)])[0]

    def leaves(self, idx) -> np.ndarray:
        idx = np.ascontiguousarray(np.array(idx, dtype=np.uint64))
        out = np.zeros((idx.size, self.num_polys), dtype=np.uint64)
        check(load().vx_batch_leaves(self._h, ptr(idx), idx.size, ptr(out)), "vx_batch_leaves")
        return out

    def prove(self, idx) -> np.ndarray:
        """merkle_tree.prove(i).siblings for every i in idx: (k, depth, 4)."""
        idx = np.ascontiguousarray(np.array(idx, dtype=np.uint64))
        depth = (self.num_leaves // self.num_caps).bit_length() - 1
        out = np.zeros((idx.size, depth, 4), dtype=np.uint64)
        check(load().vx_batch_merkle_paths(self._h, ptr(idx), idx.size, ptr(out)), "vx_batch_merkle_paths")
        return out

    def download(self, leaves: bool = True, digests: bool = True):
        """Full MerkleTree { leaves, digests } in plonky2's host layout."""
        N = self.num_leaves
        lv = np.zeros((N, self.num_polys), dtype=np.uint64) if leaves else None
        dg = np.zeros((max(2 * (N - self.num_caps), 0), 4), dtype=np.uint64) if digests else None
        check(load().vx_batch_download(self._h, ptr(lv) if leaves else None,
                                       ptr(dg) if (digests and dg.size) else None), "vx_batch_download")
        return lv, dg

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            load().vx_batch_free(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ primitives
def poseidon(states: np.ndarray, ctx: Context | None = None) -> np.ndarray:
    ctx = ctx or default_context()
    states = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, 12)
    out = np.zeros_like(states)
    check(load().vx_poseidon_permute(ctx.handle, ptr(states), states.shape[0], ptr(out)), "vx_poseidon_permute")
    return out


def hash_n_to_hash_no_pad(inputs: np.ndarray, ctx: Context | None = None) -> np.ndarray:
    """Batched: inputs (count, len) -> (count, 4)."""
    ctx = ctx or default_context()
    inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
    count, ln = inputs.shape
    out = np.zeros((count, 4), dtype=np.uint64)
    check(load().vx_hash_no_pad(ctx.handle, ptr(inputs) if ln else None, count, ln, ptr(out)), "vx_hash_no_pad")
    return out


def poseidon_bn128(states: np.ndarray, ctx: Context | None = None) -> np.ndarray:
    """`permution` (P2X/backend/wrapper/poseidon_bn128.rs:20-25), batched: states (count, 4 scalars, 4 u64 words)."""
    ctx = ctx or default_context()
    states = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, 4, 4)
    out = np.zeros_like(states)
    check(load().vx_bn128_permute(ctx.handle, ptr(states), states.shape[0], ptr(out)), "vx_bn128_permute")
    return out


def poseidon_bn128_hash(inputs: np.ndarray, or_noop: bool = False, ctx: Context | None = None) -> np.ndarray:
    """PoseidonBN128Hash::hash_no_pad / hash_or_noop (plonky2_config.rs:135-187), batched: (count, len) -> (count, 4)."""
    ctx = ctx or default_context()
    inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
    count, ln = inputs.shape
    out = np.zeros((count, 4), dtype=np.uint64)
    check(load().vx_bn128_hash(ctx.handle, ptr(inputs) if ln else None, count, ln, int(or_noop), ptr(out)),
          "vx_bn128_hash")
    return out


def poseidon_bn128_constants():
    """(C[88], S[392], M[16], P[16]) as Python ints, in the reference's layout (host-side derivation, no GPU)."""
    bufs = [np.zeros((k, 4), dtype=np.uint64) for k in (88, 392, 16, 16)]
    check(load().vx_bn128_constants(*[ptr(b) for b in bufs]), "vx_bn128_constants")
    return tuple([sum(int(w) << (64 * i) for i, w in enumerate(row)) for row in b] for b in bufs)


def poseidon_round_constants() -> np.ndarray:
    out = np.zeros(360, dtype=np.uint64)
    check(load().vx_poseidon_constants(ptr(out)), "vx_poseidon_constants")
    return out


def ntt(values: np.ndarray, inverse: bool = False, coset_shift: int = 0, ctx: Context | None = None) -> np.ndarray:
    """Natural-order (coset) NTT / iNTT of each row of a (c, n) batch: fft / ifft / coset_fft / coset_ifft."""
    ctx = ctx or default_context()
    values = np.ascontiguousarray(values, dtype=np.uint64)
    c, n = values.shape
    out = np.zeros_like(values)
    check(load().vx_ntt(ctx.handle, ptr(values), ptr(out), c, n.bit_length() - 1, int(inverse), coset_shift), "vx_ntt")
    return out
