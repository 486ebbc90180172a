"""vectorx_b200: B200-native (sm_100a) proving hot path behind VectorX's plonky2x proofs.

Public surface mirrors plonky2 v0.2.0 (`PolynomialBatch`, `MerkleTree`); everything computes in
libvectorx_b200.so through the C ABI of include/vectorx_b200.h.  No CPU fallback.
"""
from ._lib import Context, VxError, default_context, device_count, device_list, load, pinned_empty, LIB_PATH  # noqa: F401
from .plonky2 import (MerkleCap, MerkleProof, MerkleTree, PolynomialBatch, hash_n_to_hash_no_pad,  # noqa: F401
                      merkle_tree_digests, ntt, poseidon, poseidon_round_constants, reverse_bits, POSEIDON_HASH,
                      POSEIDON_BN128_HASH, poseidon_bn128, poseidon_bn128_hash, poseidon_bn128_constants)
from .prover import CircuitData, prove  # noqa: F401,E402
from .challenger import Challenger, hash_no_pad_host, poseidon_host  # noqa: F401,E402
from .local_prover import CircuitSpec, LocalProver  # noqa: F401,E402
from .proof_io import proof_from_bytes, proof_from_hex, proof_to_bytes, proof_to_hex  # noqa: F401,E402
