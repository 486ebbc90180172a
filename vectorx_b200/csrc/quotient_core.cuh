// Constraint evaluation at one LDE point: the parts shared by the bytecode interpreter (prover.cu) and the kernels that
// quotient_jit.cu compiles at run time from a circuit's gate program (NVRTC) -- parameters, point set-up, the permutation
// argument and the final division by Z_H.  Device code only, no host headers: this file is also compiled by NVRTC.
//
// Replaces plonky2 v0.2.0 plonk/vanishing_poly.rs `eval_vanishing_poly_base_batch` (reference call site:
// contracts/lib/succinctx/plonky2x/core/src/backend/circuit/build.rs:69-75).
#pragma once
#include "poseidon.cuh"
#include "twiddle_view.cuh"

struct QuotParams {
    const u64 *cs, *wires, *zpp;       // LDE column-major, leaf order, stride N
    unsigned long long N;
    unsigned int bits, rate_bits, degree_bits;
    unsigned int num_wires, num_routed, num_constants, num_selectors, num_challenges, num_pp, max_degree;
    const u64* program;
    unsigned int program_len;          // padded to a multiple of QCHUNK
    unsigned int num_regs;             // registers the program uses (shared-memory register file rows)
    const u64* beta_k;                 // [challenge][routed]  beta_k * k_j
    const u64* apow;                   // [challenge][num_terms] alpha_k^j
    unsigned int num_terms, num_perm_terms;
    u64 betas[4], gammas[4];
    u64 pi_hash[4];
    u64 zh_inv[64];                    // 1 / Z_H on the 2^rate_bits cosets of the subgroup
    u64 zh[64];
    u64 n_inv;
    u64* out;                          // [challenge][N], leaf order
    TwiddleView tw;
};

// L2 residency hints.  A point's wire (and constant / selector) columns are read again by every gate that uses them, its
// sigma and Z / partial-product columns exactly once; the columns of the points in flight (148 SMs x 768 points x 240
// columns = 218 MB) do not fit the 126 MB L2, and without hints 4.4 GB instead of the algorithmic 1.0 GB come from DRAM
// (profiles/r02_quotient_jit_ncu.md).  Read-once columns are loaded evict-first and kept out of L1, re-read ones evict-last.
GL_D u64 quot_policy_stream() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
GL_D u64 quot_policy_keep() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
GL_D u64 quot_ld_stream(const u64* a, u64 pol) {
    u64 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(a), "l"(pol));
    return v;
}
GL_D u64 quot_ld_keep(const u64* a, u64 pol) {
    u64 v;
    asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(a), "l"(pol));
    return v;
}

#define QBLOCK 128           // threads per block of the interpreter (== POSEIDON_BLOCK: the dense layer's scratch is sized for it)

struct QuotPoint {
    unsigned long long j;              // leaf (bit-reversed LDE index) this thread evaluates
    unsigned int coset;                // natural index mod 2^rate_bits: selects Z_H
    bool live;                         // idle threads shadow the last point (they take part in barriers)
};

// alpha-reduction of one constraint value into the running totals (lazy dot products, one reduction per challenge)
GL_D void quot_add_term(const QuotParams& p, GlAcc2 (&tot)[2], unsigned int idx, u64 v) {
    gl_acc2_mad(tot[0], v, __ldg(p.apow + idx));
    if (p.num_challenges > 1) gl_acc2_mad(tot[1], v, __ldg(p.apow + p.num_terms + idx));
}

// point set-up, Z(1) = 1 and the permutation argument.  The chunk loop is outermost so that every routed wire and sigma
// value is loaded ONCE and used for all challenges.
GL_D QuotPoint quot_prologue(const QuotParams& p, GlAcc2 (&tot)[2]) {
    QuotPoint q;
    const unsigned long long j_raw = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    q.live = j_raw < p.N;
    const unsigned long long j = q.live ? j_raw : p.N - 1;
    q.j = j;
    const unsigned long long N = p.N;
    const unsigned int i = (unsigned int)bitrev_u64(j, p.bits);                 // natural LDE index
    const u64 x = gl_mul_cc(GL_GENERATOR, tw_pow_view(p.tw, i << (32 - p.bits)));
    q.coset = i & ((1u << p.rate_bits) - 1);
    const u64 zh = p.zh[q.coset];
    // L_0(x) = Z_H(x) / (n (x - 1))
    const u64 l0 = gl_mul_cc(gl_mul_cc(zh, p.n_inv), gl_inv(gl_sub(x, 1)));
    const unsigned long long jn = bitrev_u64((i + (1u << p.rate_bits)) & (N - 1), p.bits);   // leaf of g_n * x
    gl_acc2_init(tot[0], 0);
    gl_acc2_init(tot[1], 0);
    const u64 pol_stream = quot_policy_stream(), pol_keep = quot_policy_keep();
    const unsigned int nch = p.num_challenges;
    const unsigned int chunks = (p.num_routed + p.max_degree - 1) / p.max_degree;
    u64 prev[2] = {0, 0};
#pragma unroll
    for (unsigned int k = 0; k < 2; k++) {
        if (k < nch) {
            const u64 z = quot_ld_stream(p.zpp + (unsigned long long)k * N + j, pol_stream);
            quot_add_term(p, tot, k, gl_mul_cc(l0, gl_sub(z, 1)));
            prev[k] = z;
        }
    }
    for (unsigned int c = 0; c < chunks; c++) {
        u64 num[2] = {1, 1}, den[2] = {1, 1};
        const unsigned int hi = min(p.num_routed, (c + 1) * p.max_degree);
        for (unsigned int w = c * p.max_degree; w < hi; w++) {
            const u64 wv = quot_ld_keep(p.wires + (unsigned long long)w * N + j, pol_keep);
            const u64 sg = quot_ld_stream(p.cs + (unsigned long long)(p.num_constants + w) * N + j, pol_stream);
#pragma unroll
            for (unsigned int k = 0; k < 2; k++) {
                if (k < nch) {
                    const u64 a = gl_add(gl_mul_add_cc(x, __ldg(p.beta_k + k * p.num_routed + w), wv), p.gammas[k]);
                    const u64 b = gl_add(gl_mul_add_cc(sg, p.betas[k], wv), p.gammas[k]);
                    num[k] = gl_mul_cc(num[k], a);
                    den[k] = gl_mul_cc(den[k], b);
                }
            }
        }
#pragma unroll
        for (unsigned int k = 0; k < 2; k++) {
            if (k < nch) {
                const u64 next = quot_ld_stream((c + 1 < chunks) ? p.zpp + (unsigned long long)(nch + k * p.num_pp + c) * N + j
                                                                 : p.zpp + (unsigned long long)k * N + jn, pol_stream);
                quot_add_term(p, tot, nch + k * chunks + c, gl_sub(gl_mul_cc(prev[k], num[k]), gl_mul_cc(next, den[k])));
                prev[k] = next;
            }
        }
    }
    return q;
}

// one partial round of the Poseidon gate in the sparse form (super-instruction PARTIAL12 of the gate bytecode)
GL_D void quot_partial12(u64 st[12], unsigned int round) {
    const u64* v = c_pos.pv + 11 * round;
    const u64* w = c_pos.pw + 11 * round;
    const u64 x0 = gl_add_canon(st[0], c_pos.pk[round]);
    GlAcc d;
    gl_acc_init(d, 0);
    gl_acc_mad_small(d, x0, 25u);
#pragma unroll
    for (int k = 1; k < 12; k++) gl_acc_mad(d, v[k - 1], st[k]);
#pragma unroll
    for (int k = 1; k < 12; k++) st[k] = gl_mul_add_cc(w[k - 1], x0, st[k]);
    st[0] = gl_acc_reduce(d);
}

// total / Z_H, canonical, leaf order
GL_D void quot_epilogue(const QuotParams& p, const QuotPoint& q, const GlAcc2 (&tot)[2]) {
    const u64 zi = p.zh_inv[q.coset];
    if (q.live) {
        p.out[q.j] = gl_canon(gl_mul_cc(gl_acc2_reduce(tot[0]), zi));
        if (p.num_challenges > 1) p.out[p.N + q.j] = gl_canon(gl_mul_cc(gl_acc2_reduce(tot[1]), zi));
    }
}
