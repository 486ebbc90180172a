// HOST stand-in for vectorx_b200/csrc/quotient_core.cuh (TEST INFRASTRUCTURE, used by tests/test_quotient_jit.py only).
// The translation unit that quotient_jit.cu generates from a gate program is compiled against this file with g++ and run
// on the CPU at single points: field operations are plain big-integer arithmetic mod p, the Poseidon super-instructions
// the textbook formulas over the tables the library exports.  What the test checks is the GENERATOR (operands, immediates,
// statement order, alpha-power indices), not the device arithmetic -- that is the GPU tests' job.
#pragma once
#include <cstring>
typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned __int128 u128;
#define GL_D static inline
#define QBLOCK 128
#define __global__
#define __shared__ static
#define __launch_bounds__(a, b)
static struct { unsigned x; } threadIdx = {0};
static inline void __syncthreads() {}
static const u64 P = 0xFFFFFFFF00000001ULL;

GL_D u64 gl_add(u64 a, u64 b) { return (u64)(((u128)(a % P) + (b % P)) % P); }
GL_D u64 gl_sub(u64 a, u64 b) { return (u64)(((u128)(a % P) + P - (b % P)) % P); }
GL_D u64 gl_mul_cc(u64 a, u64 b) { return (u64)((u128)(a % P) * (b % P) % P); }
GL_D u64 gl_mul_add_cc(u64 a, u64 b, u64 c) { return gl_add(gl_mul_cc(a, b), c); }
GL_D u64 gl_pow7_cc(u64 x) { u64 x2 = gl_mul_cc(x, x), x4 = gl_mul_cc(x2, x2); return gl_mul_cc(gl_mul_cc(x4, x2), x); }
struct GlAcc2 { u64 v; };
GL_D void gl_acc2_init(GlAcc2& t, u64 x) { t.v = x % P; }
GL_D void gl_acc2_mad(GlAcc2& t, u64 a, u64 b) { t.v = gl_add(t.v, gl_mul_cc(a, b)); }
GL_D u64 gl_acc2_reduce(const GlAcc2& t) { return t.v; }

// tables handed in by the test (vx_poseidon_constants / vx_poseidon_fast_tables of the library)
static u64 T_rc[372], T_d[144], T_e[12], T_k[22], T_v[242], T_w[242];
static struct { unsigned int rc22[31 * 36]; } c_pos;      // only its ADDRESS is used: (kl - c_pos.rc22) / 36 = round
GL_D void poseidon_mds_add_freq(u64 s[12], const unsigned int* kl) {
    static const u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    const unsigned round = (unsigned)((kl - c_pos.rc22) / 36);
    u64 out[12];
    for (int r = 0; r < 12; r++) {
        u64 acc = round < 30 ? T_rc[12 * round + r] : 0;
        for (int i = 0; i < 12; i++) acc = gl_add(acc, gl_mul_cc(s[(i + r) % 12], C[i]));
        if (r == 0) acc = gl_add(acc, gl_mul_cc(s[0], 8));
        out[r] = acc;
    }
    memcpy(s, out, sizeof out);
}
template <int STRIDE>
GL_D void poseidon_dense_layer(u64 s[12], u64*) {
    u64 out[12];
    for (int j = 0; j < 12; j++) {
        u64 acc = T_e[j];
        for (int i = 0; i < 12; i++) acc = gl_add(acc, gl_mul_cc(s[i], T_d[12 * j + i]));
        out[j] = acc;
    }
    memcpy(s, out, sizeof out);
}
GL_D void quot_partial12(u64 st[12], unsigned round) {
    const u64 x0 = gl_add(st[0], T_k[round]);
    u64 d = gl_mul_cc(x0, 25);
    for (int k = 1; k < 12; k++) d = gl_add(d, gl_mul_cc(st[k], T_v[11 * round + k - 1]));
    for (int k = 1; k < 12; k++) st[k] = gl_add(st[k], gl_mul_cc(x0, T_w[11 * round + k - 1]));
    st[0] = d;
}

struct QuotParams {
    const u64 *cs, *wires;
    unsigned long long N;
    const u64* apow;
    unsigned num_terms, num_challenges;
    u64 pi_hash[4];
    u64* out;
};
struct QuotPoint { unsigned long long j; };
GL_D QuotPoint quot_prologue(const QuotParams&, GlAcc2 (&tot)[2]) { tot[0].v = tot[1].v = 0; return QuotPoint{0}; }
GL_D void quot_epilogue(const QuotParams& p, const QuotPoint&, const GlAcc2 (&tot)[2]) { p.out[0] = tot[0].v; p.out[1] = tot[1].v; }
GL_D u64 quot_policy_keep() { return 0; }
GL_D u64 quot_ld_keep(const u64* a, u64) { return *a; }
