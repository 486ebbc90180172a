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
#define POSEIDON_PARTIAL_ROUNDS 22
#ifndef POSEIDON_BLOCK
#define POSEIDON_BLOCK 128
#endif                              // threads per block of every hashing kernel (shared scratch is sized for it)

#include "poseidon_tables.h"

// All tables (8.3 KB) live in constant memory and are only ever indexed warp-uniformly.  One copy per
// translation unit that hashes; each is filled by poseidon_upload_constants() at context creation.
#ifdef __CUDACC_RTC__
extern "C" { __constant__ PoseidonTables c_pos; }      // run-time compiled modules: filled through cudaLibraryGetGlobal
#else
static __constant__ PoseidonTables c_pos;
#endif

#ifndef __CUDACC_RTC__
static inline cudaError_t poseidon_upload_constants(const PoseidonTables& t, cudaStream_t stream) {
    cudaError_t e = cudaMemcpyToSymbolAsync(c_pos, &t, sizeof t, 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);
}
#endif

// ---- frequency-domain MDS layer --------------------------------------------------------------------------------
// The circulant part y[r] = sum_i CIRC[i] x[(r+i) % 12] is a cyclic correlation over Z/12 = Z/4 x Z/3
// (index j = 3a + 4b).  A 4-point DFT along `a` (roots +-1, +-i: additions only) turns it into three 3-point
// correlations whose constants are DFT(CIRC)/4 = {16,16,32}, {2+i, -16+i, -1+4i} (doubled, conjugate pair folded)
// and {-1,2,8}: every product is a shift.  plonky2 ships the same idea as `mds_multiply_freq` for its CPU code; the
// constants here are re-derived (tools/mds_freq_derive.py).  One call works on one 32-bit "plane" of 22-bit limbs
// with wrap-around arithmetic: intermediates may overflow 32 bits, the final value (< 2^22 * 272 + 2^22) cannot.
// k (warp-uniform) holds the constants to add, stride 3 (limb planes interleaved).
GL_D void poseidon_mds_plane(const u32 x[12], u32 y[12], const u32* __restrict__ k) {
    constexpr int G[3][4] = {{0, 3, 6, 9}, {4, 7, 10, 1}, {8, 11, 2, 5}};
    u32 X0[3], X2[3], r[3], s[3];
#pragma unroll
    for (int b = 0; b < 3; b++) {
        u32 x0 = x[G[b][0]], x1 = x[G[b][1]], x2 = x[G[b][2]], x3 = x[G[b][3]];
        u32 p = x0 + x2, q = x1 + x3;
        r[b] = x0 - x2;
        s[b] = x1 - x3;
        X0[b] = p + q;
        X2[b] = p - q;
    }
    u32 T = X0[0] + X0[1] + X0[2];
#pragma unroll
    for (int b = 0; b < 3; b++) {
        const int b1 = (b + 1) % 3, b2 = (b + 2) % 3;
        u32 Y0 = (T + X0[b2]) << 4;
        u32 Y2 = (X2[b1] << 1) + (X2[b2] << 3) - X2[b];
        u32 Y1r = (r[b] << 1) - (r[b1] << 4) - r[b2] - s[b] - s[b1] - (s[b2] << 2);
        u32 Y1i = (s[b] << 1) - (s[b1] << 4) - s[b2] + r[b] + r[b1] + (r[b2] << 2);
        u32 P = Y0 + Y2, Q = Y0 - Y2;
        y[G[b][0]] = P + Y1r + k[3 * G[b][0]];
        y[G[b][2]] = P - Y1r + k[3 * G[b][2]];
        y[G[b][1]] = Q + Y1i + k[3 * G[b][1]];
        y[G[b][3]] = Q - Y1i + k[3 * G[b][3]];
    }
    y[0] += x[0] << 3;          // DIAG[0] = 8
}

// out = MDS * s + add, add given as 22/22/20-bit limbs (kl: 12 x 3 u32, warp-uniform pointer into constant memory)
GL_D void poseidon_mds_add_freq(u64 s[12], const u32* __restrict__ kl) {
    u32 l0[12], l1[12], l2[12], y0[12], y1[12], y2[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u32 lo = lo32(s[i]), hi = hi32(s[i]);
        l0[i] = lo & 0x3fffffu;
        l1[i] = __funnelshift_r(lo, hi, 22) & 0x3fffffu;
        l2[i] = hi >> 12;
    }
    poseidon_mds_plane(l0, y0, kl);
    poseidon_mds_plane(l1, y1, kl + 1);
    poseidon_mds_plane(l2, y2, kl + 2);
#pragma unroll
    for (int r = 0; r < 12; r++) {
        // value = y0 + y1 2^22 + y2 2^44 (< 2^76)
        u64 t = mad_wide(y1[r], 1u << 22, (u64)y0[r]);
        u64 u = mad_wide(y2[r], 1u << 12, (u64)hi32(t));
        s[r] = gl_reduce96_cc(lo32(t), lo32(u), hi32(u));
    }
}

// s <- D * s + e with D a dense matrix of full-width constants (the MDS layer of full round 3 merged
// with the partial rounds' initial matrix).  Rows are produced in a rolled loop (small code) and
// staged through this thread's shared-memory column: scratch[j * POSEIDON_BLOCK].
template <int STRIDE = POSEIDON_BLOCK>
GL_D void poseidon_dense_layer(u64 s[12], u64* __restrict__ scratch, const u64* __restrict__ dense_d = c_pos.dense_d,
                               const u64* __restrict__ dense_e = c_pos.dense_e) {
#pragma unroll 1
    for (int j = 0; j < 12; j++) {
        const u64* row = dense_d + 12 * j;
        GlAcc acc;
        gl_acc_init(acc, dense_e[j]);
#pragma unroll
        for (int i = 0; i < 12; i++) gl_acc_mad(acc, row[i], s[i]);
        scratch[j * STRIDE] = gl_acc_reduce(acc);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = scratch[i * STRIDE];
}

// 22 partial rounds in the sparse form (see poseidon_tables.h), one round at a time (G = 1 of the grouped form below)
// The 22 partial rounds in groups of G: lanes 1..11 enter only linearly, so inside a group every first-row dot product is taken
// against the group's INITIAL lanes plus the cross terms pc[r][q] x_q of the S-box outputs already produced, and the lanes
// are brought up to date once per group, s_i += sum_q w[q][i] x_q -- a lazy dot product with ONE reduction per lane and group
// instead of one reduced multiply-add per lane and round.
template <int G>
GL_D void poseidon_partial_group(u64 s[12], const int r0, const u64* __restrict__ pk, const u64* __restrict__ pv,
                                 const u64* __restrict__ pw, const u64* __restrict__ pc) {
    u64 x[G];
#pragma unroll
    for (int j = 0; j < G; j++) {
        const int r = r0 + j;
        x[j] = gl_add_canon_cc(gl_pow7_cc(s[0]), pk[r]);
        GlAcc d;
        gl_acc_init(d, 0);
        gl_acc_mad_small(d, x[j], 25u);
#pragma unroll
        for (int i = 1; i < 12; i++) gl_acc_mad(d, pv[11 * r + i - 1], s[i]);
#pragma unroll
        for (int q = 0; q < j; q++) gl_acc_mad(d, pc[22 * r + r0 + q], x[q]);
        s[0] = gl_acc_reduce(d);
    }
#pragma unroll
    for (int i = 1; i < 12; i++) {
        GlAcc u;
        gl_acc_init(u, s[i]);
        gl_acc_mad_first(u, pw[11 * r0 + i - 1], x[0]);
#pragma unroll
        for (int j = 1; j < G; j++) gl_acc_mad(u, pw[11 * (r0 + j) + i - 1], x[j]);
        s[i] = gl_acc_reduce(u);
    }
}
template <int G>
GL_D void poseidon_partial_rounds_grouped(u64 s[12]) {
    constexpr int FULL = POSEIDON_PARTIAL_ROUNDS / G, TAIL = POSEIDON_PARTIAL_ROUNDS % G;
#pragma unroll 1
    for (int g = 0; g < FULL; g++) poseidon_partial_group<G>(s, g * G, c_pos.pk, c_pos.pv, c_pos.pw, c_pos.pc);
    if constexpr (TAIL > 0) poseidon_partial_group<TAIL>(s, FULL * G, c_pos.pk, c_pos.pv, c_pos.pw, c_pos.pc);
}

// Rounds per group of the partial rounds (profiles/r02_poseidon_ab.md).  Larger groups execute fewer instructions, but their
// unrolled bodies have to share the 32 KB instruction cache with the full-round body: with the warps of an SM drifting
// through different regions G = 1 / 2 / 4 / 6 / 11 hash the config-1 leaves in 8.68 / 8.13 / 9.23 / 10.85 / 10.17 ms.  G = 2 is
// the default; the long leaf sponge uses G = 4 in 256-thread blocks that re-align their warps with one barrier per
// permutation (7.99 ms) -- short kernels (one or two permutations per thread) run the big body with a cold cache and lose.
#ifndef POSEIDON_GROUP
#define POSEIDON_GROUP 2
#endif
#define POSEIDON_GROUP_LONG 4
#define POSEIDON_BLOCK_LONG 256

// scratch: this thread's column of a STRIDE-wide shared array of 12 rows
template <int G = POSEIDON_GROUP, int STRIDE = POSEIDON_BLOCK>
GL_D void poseidon_permute(u64 s[12], u64* __restrict__ scratch) {
    const u64* rc = c_pos.rc;
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_canon_cc(s[i], rc[i]);
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        const int first = half ? 27 : 1;                    // constants of the round after each full round
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = gl_pow7_cc(s[i]);
            if (half == 0 && r == 3) poseidon_dense_layer<STRIDE>(s, scratch);
            else poseidon_mds_add_freq(s, c_pos.rc22 + 36 * (first + r));
        }
        if (half == 0) {
            poseidon_partial_rounds_grouped<G>(s);
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = gl_add_canon_cc(s[i], rc[26 * 12 + i]);
        }
    }
}
