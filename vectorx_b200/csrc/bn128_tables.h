// Host-side derivation of the PoseidonBN128 tables (t = 4, x^5, 8 + 56 rounds over the BN254 scalar field).
//
// The reference ships them as 512 decimal literals (contracts/lib/succinctx/plonky2x/core/src/backend/wrapper/
// poseidon_bn128_constants.rs, indexed by poseidon_bn128.rs:27-110).  Here they are re-derived at start-up from the Poseidon
// parameter generator (Grain LFSR) and the sparse factorisation of the partial rounds; vx_bn128_constants() exports them so
// the CPU test-suite can compare them with the oracle's independent derivation (and with the reference literals when
// /root/reference is mounted).
#pragma once
#include <cstdint>

struct Bn128Fr { uint64_t l[4]; };     // little-endian 64-bit limbs

struct Bn128Tables {
    // canonical (non-Montgomery) values, reference layout: C[88], S[392], M and P row-major [j][i] as mix() indexes them
    Bn128Fr C[88], S[392], M[16], P[16];
};

// false only on an internal inconsistency (singular matrix)
bool bn128_derive_tables(Bn128Tables* out);

// host arithmetic, exposed for the table upload (Montgomery form) and tests
Bn128Fr bn128_to_mont(const Bn128Fr& a);
Bn128Fr bn128_from_mont(const Bn128Fr& a);
