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
#define POSEIDON_BLOCK 128          // threads per block of every hashing kernel (shared scratch is sized for it)

#include "poseidon_tables.h"

// All tables (8.3 KB) live in constant memory and are only ever indexed warp-uniformly.  One copy per
// translation unit that hashes; each is filled by poseidon_upload_constants() at context creation.
static __constant__ PoseidonTables c_pos;

static inline cudaError_t poseidon_upload_constants(const PoseidonTables& t, cudaStream_t stream) {
    cudaError_t e = cudaMemcpyToSymbolAsync(c_pos, &t, sizeof t, 0, cudaMemcpyHostToDevice, stream);
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
        s[r] = gl_reduce96(pack64(lo32(t), lo32(u)), hi32(u));
    }
}

// s <- D * s + e with D a dense matrix of full-width constants (the MDS layer of full round 3 merged
// with the partial rounds' initial matrix).  Rows are produced in a rolled loop (small code) and
// staged through this thread's shared-memory column: scratch[j * POSEIDON_BLOCK].
template <int ALU = 0>
GL_D void poseidon_dense_layer(u64 s[12], u64* __restrict__ scratch, const u64* __restrict__ dense_d = c_pos.dense_d,
                               const u64* __restrict__ dense_e = c_pos.dense_e) {
#pragma unroll 1
    for (int j = 0; j < 12; j++) {
        const u64* row = dense_d + 12 * j;
        GlAcc acc;
        gl_acc_init(acc, dense_e[j]);
#pragma unroll
        for (int i = 0; i < 12; i++) gl_acc_mad_v<ALU>(acc, row[i], s[i]);
        scratch[j * POSEIDON_BLOCK] = gl_acc_reduce(acc);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = scratch[i * POSEIDON_BLOCK];
}

// 22 partial rounds in the sparse form (see poseidon_tables.h)
template <int ALU = 0, int SB = 0>
GL_D void poseidon_partial_rounds(u64 s[12], const int rounds = POSEIDON_PARTIAL_ROUNDS, const u64* __restrict__ pk = c_pos.pk,
                                  const u64* __restrict__ pv = c_pos.pv, const u64* __restrict__ pw = c_pos.pw) {
#pragma unroll 1
    for (int r = 0; r < rounds; r++) {
        const u64* v = pv + 11 * r;
        const u64* w = pw + 11 * r;
        u64 x0 = gl_add_canon(gl_pow7_v<SB>(s[0]), pk[r]);
        // d = 25 * x0 + sum_i v_i s_i   (25 = MDS[0][0])
        GlAcc d;
        gl_acc_init(d, 0);
        gl_acc_mad_small(d, x0, 25u);
#pragma unroll
        for (int i = 1; i < 12; i++) gl_acc_mad_v<ALU>(d, v[i - 1], s[i]);
#pragma unroll
        for (int i = 1; i < 12; i++) s[i] = gl_mul_add_v<SB>(w[i - 1], x0, s[i]);
        s[0] = gl_acc_reduce(d);
    }
}

// scratch: this thread's column of a POSEIDON_BLOCK-wide shared array of 12 rows
// MV = 0: frequency-domain MDS layer (shifts/adds on 22-bit limb planes); MV = 1: IMAD.WIDE MDS on 32-bit halves (A/B);
// MV = 2: MV 0 with the lazy dot products accumulated on the ALU pipe (gl_acc_mad_alu, A/B)
// SB: S-box / multiply-add form (gl_pow7_v): 0 = carry-chain products (shipped), 1 / 2 = zero-extended-addend products (A/B)
template <int MV = 0, int SB = 0>
GL_D void poseidon_permute(u64 s[12], u64* __restrict__ scratch) {
    const u64* rc = c_pos.rc;
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_canon(s[i], rc[i]);
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        const int first = half ? 27 : 1;                    // constants of the round after each full round
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = gl_pow7_v<SB>(s[i]);
            if (half == 0 && r == 3) poseidon_dense_layer<MV == 2>(s, scratch);
            else if (MV != 1) poseidon_mds_add_freq(s, c_pos.rc22 + 36 * (first + r));
            else poseidon_mds_add(s, rc + 12 * (first + r));
        }
        if (half == 0) {
            poseidon_partial_rounds<MV == 2, SB>(s);
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = gl_add_canon(s[i], rc[26 * 12 + i]);
        }
    }
}

// Hybrid partial rounds (A/B): the first c_posk_naive partial rounds in the spec form -- x0^7 then the frequency-domain
// MDS layer, no IMAD.WIDE in the linear part -- and the remaining ones in the sparse form with tables derived for that
// suffix (c_posk).  naive = 22 drops the dense layer and the sparse rounds altogether.
static __constant__ PoseidonTables c_posk;
static __constant__ int c_posk_naive;

GL_D void poseidon_permute_hybrid(u64 s[12], u64* __restrict__ scratch) {
    const u64* rc = c_pos.rc;
    const int K = c_posk_naive;
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_canon(s[i], rc[i]);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_pow7_cc(s[i]);
        if (r == 3 && K == 0) poseidon_dense_layer<0>(s, scratch, c_posk.dense_d, c_posk.dense_e);
        else poseidon_mds_add_freq(s, c_pos.rc22 + 36 * (1 + r));
    }
#pragma unroll 1
    for (int j = 0; j < K; j++) {                            // spec-form partial rounds 4 .. 4 + K - 1
        s[0] = gl_pow7_cc(s[0]);
        if (j == K - 1 && K < 22) poseidon_dense_layer<0>(s, scratch, c_posk.dense_d, c_posk.dense_e);
        else poseidon_mds_add_freq(s, c_pos.rc22 + 36 * (5 + j));
    }
    if (K < 22) {
        poseidon_partial_rounds<0>(s, 22 - K, c_posk.pk, c_posk.pv, c_posk.pw);
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_add_canon(s[i], rc[26 * 12 + i]);
    }
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_pow7_cc(s[i]);
        poseidon_mds_add_freq(s, c_pos.rc22 + 36 * (27 + r));
    }
}

// =================================================================================================
// Variant S ("state in shared memory"): the same permutation with every lane loop rolled, so the
// whole hot code is a few KB and stays in the instruction cache (ncu showed ~30% of warp time in
// stall_no_instruction for the fully unrolled form), and so that a thread needs few registers.
// The state of thread t lives in its column of a 12 x POSEIDON_BLOCK shared array: lane i at
// st[i * POSEIDON_BLOCK].  The MDS layer works on 22-bit limbs with plain 32-bit IMADs
// (3 limbs x 12 terms; 22 + 9 bits of growth < 2^31) instead of IMAD.WIDE halves.
// =================================================================================================
#define PS(i) st[(i) * POSEIDON_BLOCK]

GL_D void poseidon_s_sbox_all(u64* __restrict__ st) {
#pragma unroll 1
    for (int i = 0; i < 12; i += 2) {
        u64 a = PS(i), b = PS(i + 1);
        a = gl_pow7_cc(a);
        b = gl_pow7_cc(b);
        PS(i) = a;
        PS(i + 1) = b;
    }
}

GL_D u32 imad32(u32 a, u32 b, u32 c) { return a * b + c; }

// st <- MDS * st + k,  k given as 22-bit limbs (kl: 12 x 3 u32, warp-uniform constant pointer)
GL_D void poseidon_s_mds_add(u64* __restrict__ st, const u32* __restrict__ kl) {
    constexpr u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u32 l0[12], l1[12], l2[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u64 x = PS(i);
        u32 lo = lo32(x), hi = hi32(x);
        l0[i] = lo & 0x3fffffu;
        l1[i] = __funnelshift_r(lo, hi, 22) & 0x3fffffu;
        l2[i] = hi >> 12;
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        u32 a0 = kl[3 * r], a1 = kl[3 * r + 1], a2 = kl[3 * r + 2];
#pragma unroll
        for (int i = 0; i < 12; i++) {
            a0 = imad32(l0[(i + r) % 12], C[i], a0);
            a1 = imad32(l1[(i + r) % 12], C[i], a1);
            a2 = imad32(l2[(i + r) % 12], C[i], a2);
        }
        if (r == 0) {
            a0 = imad32(l0[0], 8u, a0);
            a1 = imad32(l1[0], 8u, a1);
            a2 = imad32(l2[0], 8u, a2);
        }
        // value = a0 + a1*2^22 + a2*2^44  (< 2^76)
        u64 t = mad_wide(a1, 1u << 22, (u64)a0);
        u64 u = mad_wide(a2, 1u << 12, (u64)hi32(t));
        PS(r) = gl_reduce96(pack64(lo32(t), lo32(u)), hi32(u));
    }
}

GL_D void poseidon_s_dense_layer(u64* __restrict__ st) {
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = PS(i);
#pragma unroll 1
    for (int j = 0; j < 12; j++) {
        const u64* row = c_pos.dense_d + 12 * j;
        GlAcc acc;
        gl_acc_init(acc, c_pos.dense_e[j]);
#pragma unroll
        for (int i = 0; i < 12; i++) gl_acc_mad(acc, row[i], s[i]);
        PS(j) = gl_acc_reduce(acc);
    }
}

GL_D void poseidon_s_partial_rounds(u64* __restrict__ st) {
    u64 s0 = PS(0);
#pragma unroll 1
    for (int r = 0; r < POSEIDON_PARTIAL_ROUNDS; r++) {
        const u64* v = c_pos.pv + 11 * r;
        const u64* w = c_pos.pw + 11 * r;
        u64 x0 = gl_add_canon(gl_pow7_cc(s0), c_pos.pk[r]);
        GlAcc d;
        gl_acc_init(d, 0);
        gl_acc_mad_small(d, x0, 25u);
#pragma unroll 1
        for (int i = 1; i < 12; i++) {
            u64 si = PS(i);
            gl_acc_mad(d, v[i - 1], si);
            PS(i) = gl_mul_add_cc(w[i - 1], x0, si);
        }
        s0 = gl_acc_reduce(d);
    }
    PS(0) = s0;
}

// permutes the state held in this thread's shared-memory column
GL_D void poseidon_s_permute(u64* __restrict__ st) {
    const u64* rc = c_pos.rc;
#pragma unroll 1
    for (int i = 0; i < 12; i++) PS(i) = gl_add_canon(PS(i), rc[i]);
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        const u32* next = c_pos.rc22 + 36 * (half ? 27 : 1);
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
            poseidon_s_sbox_all(st);
            if (half == 0 && r == 3) poseidon_s_dense_layer(st);
            else poseidon_s_mds_add(st, next + 36 * r);
        }
        if (half == 0) {
            poseidon_s_partial_rounds(st);
#pragma unroll 1
            for (int i = 0; i < 12; i++) PS(i) = gl_add_canon(PS(i), rc[26 * 12 + i]);
        }
    }
}
