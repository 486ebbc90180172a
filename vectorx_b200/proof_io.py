"""Proof bytes: the wire format plonky2x uses for `ProofWithPublicInputs`.

The reference writes a proof as `"0x" + hex(bincode::serialize(&proof_with_pis))`
(contracts/lib/succinctx/plonky2x/core/src/utils/serde/mod.rs:82-96, read back at :98-111; used for
`ProofResult` at .../backend/function/result.rs and the mapreduce proof vectors).  bincode 1.x with
its default options is little-endian with fixed-width integers: a `Vec<T>` is a u64 length followed
by the elements, structs / tuples / fixed arrays are their fields in declaration order with no
framing.  The field order below is the declaration order of plonky2 v0.2.0's serde-derived structs
(plonk/proof.rs `ProofWithPublicInputs`, `Proof`, `OpeningSet`; fri/proof.rs `FriProof`,
`FriQueryRound`, `FriInitialTreeProof`, `FriQueryStep`; hash/merkle_proofs.rs `MerkleProof`;
hash/merkle_tree.rs `MerkleCap`; hash/hash_types.rs `HashOut`), restated from the published source
because the crate is not vendored in the reference tree -- the reference holds no golden proof bytes,
so this layout is PARITY UNPINNED:

  ProofWithPublicInputs { proof, public_inputs: Vec<F> }
  Proof        { wires_cap, plonk_zs_partial_products_cap, quotient_polys_cap: MerkleCap,
                 openings: OpeningSet, opening_proof: FriProof }
  MerkleCap    ( Vec<HashOut> )                     HashOut { elements: [F; 4] }
  OpeningSet   { constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next, partial_products,
                 quotient_polys, lookup_zs, lookup_zs_next: Vec<[F; 2]> }
  FriProof     { commit_phase_merkle_caps: Vec<MerkleCap>, query_round_proofs: Vec<FriQueryRound>,
                 final_poly: PolynomialCoeffs { coeffs: Vec<[F; 2]> }, pow_witness: F }
  FriQueryRound{ initial_trees_proof: FriInitialTreeProof { evals_proofs: Vec<(Vec<F>, MerkleProof)> },
                 steps: Vec<FriQueryStep { evals: Vec<[F; 2]>, merkle_proof: MerkleProof }> }
  MerkleProof  { siblings: Vec<HashOut> }
  F = GoldilocksField(u64): serde writes the raw u64 (this library only emits canonical values).

`x_index` is not part of a proof (the verifier re-derives it from the transcript); `proof_from_bytes`
returns it as None.
"""
from __future__ import annotations

import numpy as np

OPENING_FIELDS = ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products",
                  "quotient_polys", "lookup_zs", "lookup_zs_next")


def _u64s(out: list, values) -> None:
    a = np.ascontiguousarray(np.asarray(values, dtype=np.uint64)).reshape(-1)
    out.append(a.astype("<u8", copy=False).tobytes())


def _len(out: list, n: int) -> None:
    out.append(int(n).to_bytes(8, "little"))


def _hashes(out: list, hashes) -> None:
    """Vec<HashOut>: length, then 4 u64 per hash."""
    h = np.asarray(hashes, dtype=np.uint64).reshape(-1, 4)
    _len(out, h.shape[0])
    _u64s(out, h)


def _ext_vec(out: list, elems) -> None:
    """Vec<QuadraticExtension>: length, then 2 u64 per element.  Accepts [a, b] pairs or objects with .a / .b."""
    _len(out, len(elems))
    flat = []
    for e in elems:
        if hasattr(e, "a"):
            flat += [int(e.a), int(e.b)]
        else:
            flat += [int(e[0]), int(e[1])]
    _u64s(out, flat)


def proof_to_bytes(proof: dict) -> bytes:
    """bincode::serialize(&ProofWithPublicInputs) for a proof dict as returned by `prove`."""
    out: list[bytes] = []
    for key in ("wires_cap", "zs_pp_cap", "quotient_cap"):
        _hashes(out, proof[key])
    op = proof["openings"]
    for key in OPENING_FIELDS:
        _ext_vec(out, op.get(key, []))
    _len(out, len(proof["fri_caps"]))
    for cap in proof["fri_caps"]:
        _hashes(out, cap)
    _len(out, len(proof["queries"]))
    for q in proof["queries"]:
        _len(out, len(q["initial"]))
        for row, path in q["initial"]:
            row = np.asarray(row, dtype=np.uint64).reshape(-1)
            _len(out, row.size)
            _u64s(out, row)
            _hashes(out, path)
        _len(out, len(q["steps"]))
        for evals, path in q["steps"]:
            ev = np.asarray(evals, dtype=np.uint64).reshape(-1, 2)
            _len(out, ev.shape[0])
            _u64s(out, ev)
            _hashes(out, path)
    _ext_vec(out, proof["final_poly"])
    _u64s(out, [int(proof["pow_witness"])])
    _len(out, len(proof["public_inputs"]))
    _u64s(out, [int(x) for x in proof["public_inputs"]])
    return b"".join(out)


def proof_to_hex(proof: dict) -> str:
    """The string `serialize_proof_with_pis` writes (serde/mod.rs:82-96)."""
    return "0x" + proof_to_bytes(proof).hex()


class _Reader:
    def __init__(self, data: bytes):
        self.a = np.frombuffer(data, dtype="<u8") if len(data) % 8 == 0 else None
        if self.a is None:
            raise ValueError(f"proof bytes: length {len(data)} is not a multiple of 8")
        self.pos = 0

    def u64(self) -> int:
        if self.pos >= self.a.size:
            raise ValueError("proof bytes: truncated")
        v = int(self.a[self.pos])
        self.pos += 1
        return v

    def take(self, n: int) -> np.ndarray:
        if n < 0 or self.pos + n > self.a.size:
            raise ValueError("proof bytes: truncated")
        v = self.a[self.pos:self.pos + n].astype(np.uint64)
        self.pos += n
        return v

    def hashes(self) -> np.ndarray:
        return self.take(4 * self.u64()).reshape(-1, 4)

    def ext_vec(self) -> list:
        return [[int(a), int(b)] for a, b in self.take(2 * self.u64()).reshape(-1, 2)]


def proof_from_bytes(data: bytes) -> dict:
    """Inverse of `proof_to_bytes` (bincode::deserialize).  Raises ValueError on truncated or trailing bytes."""
    r = _Reader(bytes(data))
    proof = {"wires_cap": r.hashes(), "zs_pp_cap": r.hashes(), "quotient_cap": r.hashes()}
    proof["openings"] = {key: r.ext_vec() for key in OPENING_FIELDS}
    proof["fri_caps"] = [r.hashes() for _ in range(r.u64())]
    queries = []
    for _ in range(r.u64()):
        initial = []
        for _ in range(r.u64()):
            row = r.take(r.u64())
            initial.append((row, r.hashes()))
        steps = []
        for _ in range(r.u64()):
            evals = r.take(2 * r.u64())
            steps.append((evals, r.hashes()))
        queries.append({"x_index": None, "initial": initial, "steps": steps})
    proof["queries"] = queries
    proof["final_poly"] = r.ext_vec()
    proof["pow_witness"] = r.u64()
    proof["public_inputs"] = [int(x) for x in r.take(r.u64())]
    if r.pos != r.a.size:
        raise ValueError(f"proof bytes: {8 * (r.a.size - r.pos)} trailing bytes")
    return proof


def proof_from_hex(s: str) -> dict:
    """`deserialize_proof_with_pis` (serde/mod.rs:98-111)."""
    if not s.startswith("0x"):
        raise ValueError("proof hex: missing 0x prefix")
    return proof_from_bytes(bytes.fromhex(s[2:]))
