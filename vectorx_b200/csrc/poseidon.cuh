// Width-12 Poseidon permutation over Goldilocks, one permutation per thread.
//
// Replaces plonky2 v0.2.0 hash/poseidon.rs `Poseidon::poseidon` + poseidon_goldilocks.rs (not
// vendored in the reference; pinned by the reference KAT at
// contracts/lib/succinctx/plonky2x/core/src/frontend/hash/poseidon/poseidon256.rs:163-202).
//
// Round structure is the spec form: 4 full + 22 partial + 4 full rounds, each
//   state += RC[r];  S-box x^7 (all lanes / lane 0);  state = MDS * state
// with MDS row r = circulant(17,15,41,16,2,28,13,13,39,18,34,20) + diag(8,0,...,0).
//
// sm_100a mapping: the MDS layer is evaluated on 32-bit halves with 64-bit IMAD.WIDE.U32
// accumulators (coefficients sum to 264 < 2^9 so no carries), and the NEXT round's constants seed
// those accumulators, so the constant layer costs no instructions of its own.  Round constants live
// in __constant__ memory and are read with warp-uniform indices (LDC / constant-bank operands).
#pragma once
#include "gl.cuh"

#define POSEIDON_WIDTH 12
#define POSEIDON_RATE 8
#define POSEIDON_ROUNDS 30

// RC[30][12] followed by one all-zero row (seed for the last MDS layer).  One copy per translation
// unit that hashes; each is filled by poseidon_upload_constants() at context creation.
static __constant__ u64 c_poseidon_rc[(POSEIDON_ROUNDS + 1) * POSEIDON_WIDTH];

static inline cudaError_t poseidon_upload_constants(const u64 rc360[360], cudaStream_t stream) {
    u64 rc[(POSEIDON_ROUNDS + 1) * POSEIDON_WIDTH] = {0};
    for (int i = 0; i < 360; i++) rc[i] = rc360[i];
    cudaError_t e = cudaMemcpyToSymbolAsync(c_poseidon_rc, rc, sizeof rc, 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);
}

// out = MDS * s + add   (add = 12 canonical constants, warp-uniform pointer into constant memory)
GL_D void poseidon_mds_add(u64 s[12], const u64* __restrict__ add) {
    constexpr u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u32 lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        lo[i] = lo32(s[i]);
        hi[i] = hi32(s[i]);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        u64 k = add[r];
        u64 al = (u64)lo32(k), ah = (u64)hi32(k);
#pragma unroll
        for (int i = 0; i < 12; i++) {
            al = mad_wide(lo[(i + r) % 12], C[i], al);
            ah = mad_wide(hi[(i + r) % 12], C[i], ah);
        }
        if (r == 0) {
            al = mad_wide(lo[0], 8u, al);
            ah = mad_wide(hi[0], 8u, ah);
        }
        // value = al + ah * 2^32, al, ah < 2^42
        u64 l = al + ((u64)lo32(ah) << 32);
        u32 c = l < al;
        s[r] = gl_reduce96(l, hi32(ah) + c);
    }
}

GL_D void poseidon_permute(u64 s[12]) {
    const u64* rc = c_poseidon_rc;
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_canon(s[i], rc[i]);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_pow7(s[i]);
        poseidon_mds_add(s, rc + 12 * (r + 1));
    }
#pragma unroll 1
    for (int r = 4; r < 26; r++) {
        s[0] = gl_pow7(s[0]);
        poseidon_mds_add(s, rc + 12 * (r + 1));
    }
#pragma unroll 1
    for (int r = 26; r < 30; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_pow7(s[i]);
        poseidon_mds_add(s, rc + 12 * (r + 1));
    }
}
