// Goldilocks field arithmetic for sm_100a, p = 2^64 - 2^32 + 1.
//
// Replaces plonky2_field v0.2.0 goldilocks_field.rs (type bound in the reference at
// contracts/lib/succinctx/plonky2x/core/src/backend/circuit/config.rs:37).  Values travel as u64
// holding ANY representative (plonky2 does the same); gl_canon() produces the canonical one and
// every kernel canonicalises what it stores to HBM.
//
// The SM has no 64-bit integer datapath: everything below is written so that ptxas emits
// IMAD.WIDE.U32 (32x32+64 -> 64 on the FMA pipe) for products and IADD3/.X chains on the ALU pipe
// for the reductions, keeping both pipes busy.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

typedef unsigned long long u64;
typedef unsigned int u32;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL
#define GL_GENERATOR 14293326489335486720ULL        // MULTIPLICATIVE_GROUP_GENERATOR = coset_shift()
#define GL_POWER_OF_TWO_GENERATOR 7277203076849721926ULL   // order 2^32

#define GL_HD __host__ __device__ __forceinline__
#define GL_D __device__ __forceinline__

GL_D u32 lo32(u64 x) { return (u32)x; }
GL_D u32 hi32(u64 x) { return (u32)(x >> 32); }
GL_D u64 pack64(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }

GL_D u64 mad_wide(u32 a, u32 b, u64 c) {     // a*b + c, caller guarantees no 64-bit overflow
    u64 d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
GL_D u64 mul_wide(u32 a, u32 b) {
    u64 d;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b));
    return d;
}

GL_D u64 gl_canon(u64 a) { return a >= GL_P ? a - GL_P : a; }

// a + b where b is canonical (< p): single conditional fix-up cannot overflow twice.
GL_D u64 gl_add_canon(u64 a, u64 b) {
    u64 s = a + b;
    return s < a ? s + GL_EPS : s;
}
// general a + b (any representatives)
GL_D u64 gl_add(u64 a, u64 b) { return gl_add_canon(a, gl_canon(b)); }
// a - b (any representatives)
GL_D u64 gl_sub(u64 a, u64 b) {
    u64 bc = gl_canon(b);
    u64 d = a - bc;
    return a < bc ? d - GL_EPS : d;
}
GL_D u64 gl_neg(u64 a) {
    u64 c = gl_canon(a);
    return c ? GL_P - c : 0;
}

// x = hi*2^64 + lo  ->  x mod p as some u64 representative.
// 2^64 = 2^32 - 1, 2^96 = -1 (mod p):  x = lo - hi_hi + hi_lo*(2^32-1).
GL_D u64 gl_reduce128(u64 lo, u64 hi) {
    u32 hh = hi32(hi), hl = lo32(hi);
    u64 t = lo - hh;
    if (lo < (u64)hh) t -= GL_EPS;          // borrow: add p == subtract eps mod 2^64
    u64 m = mul_wide(hl, 0xFFFFFFFFu);
    u64 r = t + m;
    if (r < t) r += GL_EPS;                 // carry: 2^64 = eps
    return r;
}

// x = hi32*2^64 + lo (a 96-bit value): cheaper tail for small-constant accumulations
GL_D u64 gl_reduce96(u64 lo, u32 hi) {
    u64 m = mul_wide(hi, 0xFFFFFFFFu);
    u64 r = lo + m;
    if (r < lo) r += GL_EPS;
    return r;
}

// ---- carry-flag versions (PTX mad.cc / add.cc chains).  ptxas fuses each mad.lo.cc + madc.hi.cc pair into one
// IMAD.WIDE.U32 with a carry-out predicate, and the fix-ups come straight from the carry flag.
//
// eps as an opaque run-time constant: with the literal 0xffffffff ptxas splits x*eps + t into IMAD.HI (4 issue cycles on
// the FMA-heavy pipe) + IMAD.IADD; from a constant-bank operand it stays ONE IMAD.WIDE.U32 with carry-out.
static __constant__ u32 c_gl_eps = 0xFFFFFFFFu;

// x = x3*2^96 + x2*2^64 + x1*2^32 + x0  ->  some u64 representative of x mod p.
//   t = x2*eps + (x1:x0)  (carry C),  t -= x3  (borrow B);  x = t + (C - B)*2^64 = t + (C - B)*eps  (mod p)
// and t + (C - B)*eps always lands in [0, 2^64): C = 1 needs t <= 2^64 - 2^33 + 1, B = 1 (x3 < 2^32) needs
// t >= 2^64 - 2^32 + 1, so the single signed fix-up can neither overflow nor underflow.
// x2*eps + (x1:x0) is ONE accumulating IMAD.WIDE.U32 with carry-out (1 FMA-heavy + 7 ALU instructions: C - B comes out of
// ONE IADD3.X -- `addc w2, -1, 0` after the multiply-add and `addc w2, w2, 0` after the subtraction, whose CC.CF is the hardware
// carry = 1 - B, are two carries into the same register, which ptxas fuses).  Measured and
// rejected (profiles/r01*, profiles/r02_poseidon_ab.md): x2*eps = (x2 << 32) - x2 on the ALU pipe (12 ALU instructions, leaf
// hashing 9.56 vs 8.64 ms), and the two fix-ups as predicated +-eps adds (ptxas materialises the predicates: more code).
GL_D u64 gl_reduce128_cc(u32 x0, u32 x1, u32 x2, u32 x3) {
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 t0, t1, w2, s;\n\t"
        "mad.lo.cc.u32 t0, %4, %6, %2;\n\t"
        "madc.hi.cc.u32 t1, %4, %6, %3;\n\t"
        "addc.u32 w2, 0xffffffff, 0;\n\t"        // -1 + C
        "sub.cc.u32 t0, t0, %5;\n\t"
        "subc.cc.u32 t1, t1, 0;\n\t"
        "addc.u32 w2, w2, 0;\n\t"                // + (1 - B): CC.CF after a subtraction is the hardware carry = no borrow
        "shr.s32 s, w2, 31;\n\t"
        "sub.cc.u32 %0, t0, w2;\n\t"              // t += w2*eps = (w2 << 32) - sext(w2)
        "subc.u32 t1, t1, s;\n\t"
        "add.u32 %1, t1, w2;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(c_gl_eps));
    return pack64(r0, r1);
}

// x = hi*2^64 + (x1:x0), hi*eps + (x1:x0) < 2^65 - 2^33: the carry of the accumulating IMAD.WIDE is the only fix-up and
// r + eps cannot carry again (r < 2^64 - 2^33 after a wrap).  4 instructions; the compare-and-select form of
// gl_reduce96 costs 8 (two ISETP, a 64-bit add, two SEL).
GL_D u64 gl_reduce96_cc(u32 x0, u32 x1, u32 hi) {
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 t0, t1, m;\n\t"
        "mad.lo.cc.u32 t0, %4, %5, %2;\n\t"
        "madc.hi.cc.u32 t1, %4, %5, %3;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, t0, m;\n\t"             // t += C eps = (C << 32) - C
        "subc.u32 t1, t1, 0;\n\t"
        "add.u32 %1, t1, m;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(x0), "r"(x1), "r"(hi), "r"(c_gl_eps));
    return pack64(r0, r1);
}
// a + b, b canonical (< p), on the carry flag: 5 instructions against 8 for the compare-and-select form
GL_D u64 gl_add_canon_cc(u64 a, u64 b) {
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 s0, s1, m;\n\t"
        "add.cc.u32 s0, %2, %4;\n\t"
        "addc.cc.u32 s1, %3, %5;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, s0, m;\n\t"             // s += C eps: s < p - 1 after a wrap, cannot carry again
        "subc.u32 s1, s1, 0;\n\t"
        "add.u32 %1, s1, m;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
    return pack64(r0, r1);
}

// 64 x 64 -> 128: four IMAD.WIDE (one carrying out) and one three-word carry chain
GL_D void gl_mul128_cc(u64 a, u64 b, u32& r0, u32& r1, u32& r2, u32& r3) {
    u32 a0 = lo32(a), a1 = hi32(a), b0 = lo32(b), b1 = hi32(b);
    asm("{\n\t"
        ".reg .u64 p0, p1, p3;\n\t"
        ".reg .u32 h0, l1, h1, m0, m1, m2, l3, h3;\n\t"
        "mul.wide.u32 p0, %4, %6;\n\t"
        "mul.wide.u32 p1, %4, %7;\n\t"
        "mul.wide.u32 p3, %5, %7;\n\t"
        "mov.b64 {%0, h0}, p0;\n\t"
        "mov.b64 {l1, h1}, p1;\n\t"
        "mov.b64 {l3, h3}, p3;\n\t"
        "mad.lo.cc.u32 m0, %5, %6, l1;\n\t"       // (m2:m1:m0) = a1*b0 + a0*b1
        "madc.hi.cc.u32 m1, %5, %6, h1;\n\t"
        "addc.u32 m2, h3, 0;\n\t"               // two addc into one register: ptxas emits ONE IADD3.X with two carry-ins
        "add.cc.u32 %1, h0, m0;\n\t"
        "addc.cc.u32 %2, l3, m1;\n\t"
        "addc.u32 %3, m2, 0;\n\t"
        "}"
        : "=r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}

GL_D u64 gl_mul_cc(u64 a, u64 b) {
    u32 r0, r1, r2, r3;
    gl_mul128_cc(a, b, r0, r1, r2, r3);
    return gl_reduce128_cc(r0, r1, r2, r3);
}

// a * b + c (all u64, any representatives) reduced.  The addend rides in the accumulators of the two IMAD.WIDE
// whose products leave room for a 32-bit addend: (2^32-1)^2 + 2^32-1 < 2^64.
GL_D u64 gl_mul_add_cc(u64 a, u64 b, u64 c) {
    u32 a0 = lo32(a), a1 = hi32(a), b0 = lo32(b), b1 = hi32(b);
    u64 p0 = mad_wide(a0, b0, (u64)lo32(c));
    u64 p1 = mad_wide(a0, b1, (u64)hi32(c));
    u64 p3 = mul_wide(a1, b1);
    u32 r1, r2, r3;
    asm("{\n\t"
        ".reg .u32 m0, m1, m2;\n\t"
        "mad.lo.cc.u32 m0, %3, %4, %5;\n\t"
        "madc.hi.cc.u32 m1, %3, %4, %6;\n\t"
        "addc.u32 m2, %9, 0;\n\t"
        "add.cc.u32 %0, %7, m0;\n\t"
        "addc.cc.u32 %1, %8, m1;\n\t"
        "addc.u32 %2, m2, 0;\n\t"
        "}"
        : "=&r"(r1), "=&r"(r2), "=&r"(r3)
        : "r"(a1), "r"(b0), "r"(lo32(p1)), "r"(hi32(p1)), "r"(hi32(p0)), "r"(lo32(p3)), "r"(hi32(p3)));
    return gl_reduce128_cc(lo32(p0), r1, r2, r3);
}

// a^2: three products, the doubled cross term as a second accumulating IMAD.WIDE (FMA pipe) rather than a 64-bit doubling
// on the ALU pipe (2 ALU instructions fewer per squaring; leaf hashing 8.68 -> 8.59 ms, profiles/r02_poseidon_ab.md)
GL_D u64 gl_sqr_cc(u64 a) {
    u32 a0 = lo32(a), a1 = hi32(a);
    u32 r0, r1, r2, r3;
    asm("{\n\t"
        ".reg .u64 p0, p1, p3;\n\t"
        ".reg .u32 h0, l1, h1, m0, m1, m2, l3, h3;\n\t"
        "mul.wide.u32 p0, %4, %4;\n\t"
        "mul.wide.u32 p1, %4, %5;\n\t"
        "mul.wide.u32 p3, %5, %5;\n\t"
        "mov.b64 {%0, h0}, p0;\n\t"
        "mov.b64 {l1, h1}, p1;\n\t"
        "mov.b64 {l3, h3}, p3;\n\t"
        "mad.lo.cc.u32 m0, %4, %5, l1;\n\t"       // (m2:m1:m0) = 2 * a0*a1
        "madc.hi.cc.u32 m1, %4, %5, h1;\n\t"
        "addc.u32 m2, h3, 0;\n\t"               // two addc into one register: ptxas emits ONE IADD3.X with two carry-ins
        "add.cc.u32 %1, h0, m0;\n\t"
        "addc.cc.u32 %2, l3, m1;\n\t"
        "addc.u32 %3, m2, 0;\n\t"
        "}"
        : "=r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3)
        : "r"(a0), "r"(a1));
    return gl_reduce128_cc(r0, r1, r2, r3);
}

GL_D u64 gl_pow7_cc(u64 x) {
    u64 x2 = gl_sqr_cc(x);
    u64 x4 = gl_sqr_cc(x2);
    u64 x3 = gl_mul_cc(x2, x);
    return gl_mul_cc(x3, x4);
}

// a + b and a - b for ANY u64 representatives, as pure carry chains (8 ALU instructions, no compare / select):
//   a + b = s + C 2^64 = s + C eps (mod p); the fix-up itself can carry once more (only when s >= p), and then the
//   second fix-up lands below 2 eps.  Same with borrows for the difference (-2^64 = -eps).
GL_D u64 gl_add_cc(u64 a, u64 b) {
    // ptxas keeps CC.CF as the HARDWARE carry (a borrow is CF = 0), so "subc m, 0, 0" after an add chain gives C - 1,
    // not -C: one NOT turns it into the mask C eps.  (The IMAD.WIDE form c * eps + s measured slower: 1.56 vs 1.49 ms LDE.)
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 s0, s1, m;\n\t"
        "add.cc.u32 s0, %2, %4;\n\t"
        "addc.cc.u32 s1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"                  // C - 1
        "not.b32 m, m;\n\t"                      // -C = C eps (low word)
        "add.cc.u32 s0, s0, m;\n\t"              // s += C eps, carry C2 (only when s >= p)
        "addc.cc.u32 s1, s1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "not.b32 m, m;\n\t"
        "add.cc.u32 %0, s0, m;\n\t"              // s += C2 eps: s < eps here, cannot carry
        "addc.u32 %1, s1, 0;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
    return pack64(r0, r1);
}
GL_D u64 gl_sub_cc(u64 a, u64 b) {
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 d0, d1, m;\n\t"
        "sub.cc.u32 d0, %2, %4;\n\t"
        "subc.cc.u32 d1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"                  // m = -B
        "sub.cc.u32 d0, d0, m;\n\t"              // d -= B eps
        "subc.cc.u32 d1, d1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, d0, m;\n\t"
        "subc.u32 %1, d1, 0;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
    return pack64(r0, r1);
}

// ---- lazy dot products: sum_i a_i * b_i accumulated UNREDUCED in three column accumulators
//   T0 += a0*b0,  T1 += a0*b1 + a1*b0,  T2 += a1*b1      (value = T0 + T1*2^32 + T2*2^64)
// each a 64-bit IMAD.WIDE accumulator plus a carry counter, so a term costs 4 multiply-adds and 4
// carry adds and the whole sum is reduced once (a reduced multiply costs ~3x that on the ALU pipe).
struct GlAcc {
    u32 l0, h0, c0, l1, h1, c1, l2, h2, c2;
};
GL_D void gl_acc_init(GlAcc& t, u64 init) {
    t.l0 = lo32(init); t.h0 = hi32(init);
    t.c0 = t.l1 = t.h1 = t.c1 = t.l2 = t.h2 = t.c2 = 0;
}
GL_D void gl_acc_mad(GlAcc& t, u64 a, u64 b) {
    u32 a0 = lo32(a), a1 = hi32(a), b0 = lo32(b), b1 = hi32(b);
    asm("{\n\t"
        "mad.lo.cc.u32 %0, %9, %11, %0;\n\t"  "madc.hi.cc.u32 %1, %9, %11, %1;\n\t"  "addc.u32 %2, %2, 0;\n\t"
        "mad.lo.cc.u32 %3, %9, %12, %3;\n\t"  "madc.hi.cc.u32 %4, %9, %12, %4;\n\t"  "addc.u32 %5, %5, 0;\n\t"
        "mad.lo.cc.u32 %3, %10, %11, %3;\n\t" "madc.hi.cc.u32 %4, %10, %11, %4;\n\t" "addc.u32 %5, %5, 0;\n\t"
        "mad.lo.cc.u32 %6, %10, %12, %6;\n\t" "madc.hi.cc.u32 %7, %10, %12, %7;\n\t" "addc.u32 %8, %8, 0;\n\t"
        "}"
        : "+r"(t.l0), "+r"(t.h0), "+r"(t.c0), "+r"(t.l1), "+r"(t.h1), "+r"(t.c1), "+r"(t.l2), "+r"(t.h2), "+r"(t.c2)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}
// First term onto an accumulator that holds only an initial value in its T0 column (gl_acc_init): products onto the empty
// T1 / T2 columns cannot carry, which ptxas cannot know (it emits a carry-out predicate and a SEL for each).
GL_D void gl_acc_mad_first(GlAcc& t, u64 a, u64 b) {
    u32 a0 = lo32(a), a1 = hi32(a), b0 = lo32(b), b1 = hi32(b);
    asm("{\n\t"
        ".reg .u64 p1, p2;\n\t"
        ".reg .u32 q0, q1;\n\t"
        "mul.wide.u32 p1, %8, %11;\n\t"
        "mul.wide.u32 p2, %9, %11;\n\t"
        "mov.b64 {q0, q1}, p1;\n\t"
        "mov.b64 {%6, %7}, p2;\n\t"
        "mad.lo.cc.u32 %0, %8, %10, %0;\n\t"  "madc.hi.cc.u32 %1, %8, %10, %1;\n\t"  "addc.u32 %2, 0, 0;\n\t"
        "mad.lo.cc.u32 %3, %9, %10, q0;\n\t"  "madc.hi.cc.u32 %4, %9, %10, q1;\n\t"  "addc.u32 %5, 0, 0;\n\t"
        "}"
        : "+r"(t.l0), "+r"(t.h0), "=r"(t.c0), "=r"(t.l1), "=r"(t.h1), "=r"(t.c1), "=r"(t.l2), "=r"(t.h2)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    t.c2 = 0;
}
// small-constant term: a * k, k < 2^32
GL_D void gl_acc_mad_small(GlAcc& t, u64 a, u32 k) {
    u32 a0 = lo32(a), a1 = hi32(a);
    asm("{\n\t"
        "mad.lo.cc.u32 %0, %6, %8, %0;\n\t"  "madc.hi.cc.u32 %1, %6, %8, %1;\n\t"  "addc.u32 %2, %2, 0;\n\t"
        "mad.lo.cc.u32 %3, %7, %8, %3;\n\t"  "madc.hi.cc.u32 %4, %7, %8, %4;\n\t"  "addc.u32 %5, %5, 0;\n\t"
        "}"
        : "+r"(t.l0), "+r"(t.h0), "+r"(t.c0), "+r"(t.l1), "+r"(t.h1), "+r"(t.c1)
        : "r"(a0), "r"(a1), "r"(k));
}
// value = w0 + w1 2^32 + w2 2^64 + w3 2^96 + w4 2^128 = (w1:w0) - (w4:w3) + w2*eps  (2^96 = -1, 2^128 = -2^32)
GL_D u64 gl_acc_reduce(const GlAcc& t) {
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 w1, w2, w3, w4, t0, t1, m, c;\n\t"
        "add.cc.u32 w1, %3, %5;\n\t"            // w1 = h0 + l1
        "addc.cc.u32 w2, %4, %6;\n\t"           // w2 = c0 + h1 + cy
        "addc.cc.u32 w3, %7, %9;\n\t"           // w3 = c1 + h2 + cy
        "addc.u32 w4, %10, 0;\n\t"              // w4 = c2 + cy
        "add.cc.u32 w2, w2, %8;\n\t"            // w2 += l2
        "addc.cc.u32 w3, w3, 0;\n\t"
        "addc.u32 w4, w4, 0;\n\t"
        "mad.lo.cc.u32 t0, w2, %11, %2;\n\t"     // t = w2*eps + (w1:w0), carry C
        "madc.hi.cc.u32 t1, w2, %11, w1;\n\t"
        "addc.u32 c, 0xffffffff, 0;\n\t"       // -1 + C
        "sub.cc.u32 t0, t0, w3;\n\t"            // t -= (w4:w3): hardware carry = 1 - B
        "subc.cc.u32 t1, t1, w4;\n\t"
        "addc.u32 c, c, 0;\n\t"                 // c = C - B, both carries taken by ONE IADD3.X
        "shr.s32 m, c, 31;\n\t"
        "sub.cc.u32 %0, t0, c;\n\t"             // t += c*eps
        "subc.u32 t1, t1, m;\n\t"
        "add.u32 %1, t1, c;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(t.l0), "r"(t.h0), "r"(t.c0), "r"(t.l1), "r"(t.h1), "r"(t.c1), "r"(t.l2), "r"(t.h2), "r"(t.c2), "r"(c_gl_eps));
    return pack64(r0, r1);
}

// The same accumulator with each 64-bit column held as ONE 64-bit register (an aligned pair): for accumulators that live
// across a long loop (the quotient interpreter's alpha sums) the register allocator otherwise parks the halves in
// unpaired registers and moves them in and out of IMAD.WIDE pairs around every use.
struct GlAcc2 {
    u64 t0, t1, t2;
    u32 c0, c1, c2;
};
GL_D void gl_acc2_init(GlAcc2& t, u64 init) { t.t0 = init; t.t1 = t.t2 = 0; t.c0 = t.c1 = t.c2 = 0; }
GL_D void gl_acc2_mad(GlAcc2& t, u64 a, u64 b) {
    u32 a0 = lo32(a), a1 = hi32(a), b0 = lo32(b), b1 = hi32(b);
    asm("{\n\t"
        ".reg .u32 l, h;\n\t"
        "mov.b64 {l, h}, %0;\n\t"
        "mad.lo.cc.u32 l, %6, %8, l;\n\t"  "madc.hi.cc.u32 h, %6, %8, h;\n\t"  "addc.u32 %3, %3, 0;\n\t"
        "mov.b64 %0, {l, h};\n\t"
        "mov.b64 {l, h}, %1;\n\t"
        "mad.lo.cc.u32 l, %6, %9, l;\n\t"  "madc.hi.cc.u32 h, %6, %9, h;\n\t"  "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 l, %7, %8, l;\n\t"  "madc.hi.cc.u32 h, %7, %8, h;\n\t"  "addc.u32 %4, %4, 0;\n\t"
        "mov.b64 %1, {l, h};\n\t"
        "mov.b64 {l, h}, %2;\n\t"
        "mad.lo.cc.u32 l, %7, %9, l;\n\t"  "madc.hi.cc.u32 h, %7, %9, h;\n\t"  "addc.u32 %5, %5, 0;\n\t"
        "mov.b64 %2, {l, h};\n\t"
        "}"
        : "+l"(t.t0), "+l"(t.t1), "+l"(t.t2), "+r"(t.c0), "+r"(t.c1), "+r"(t.c2)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}
GL_D u64 gl_acc2_reduce(const GlAcc2& t) {
    GlAcc g;
    g.l0 = lo32(t.t0); g.h0 = hi32(t.t0); g.c0 = t.c0;
    g.l1 = lo32(t.t1); g.h1 = hi32(t.t1); g.c1 = t.c1;
    g.l2 = lo32(t.t2); g.h2 = hi32(t.t2); g.c2 = t.c2;
    return gl_acc_reduce(g);
}

// full 64x64 -> 128 product on IMAD.WIDE.U32: 4 multiplies, no carry chains
// (each partial sum provably fits 64 bits).
GL_D void gl_mul128(u64 a, u64 b, u64& lo, u64& hi) {
    u32 a0 = lo32(a), a1 = hi32(a), b0 = lo32(b), b1 = hi32(b);
    u64 p00 = mul_wide(a0, b0);
    u64 mid = mad_wide(a0, b1, (u64)hi32(p00));
    u64 mid2 = mad_wide(a1, b0, (u64)lo32(mid));
    u64 top = mad_wide(a1, b1, (u64)hi32(mid));
    hi = top + hi32(mid2);
    lo = pack64(lo32(p00), lo32(mid2));
}

GL_D u64 gl_mul(u64 a, u64 b) {
    u64 lo, hi;
    gl_mul128(a, b, lo, hi);
    return gl_reduce128(lo, hi);
}

GL_D u64 gl_sqr(u64 a) {
    u32 a0 = lo32(a), a1 = hi32(a);
    u64 p00 = mul_wide(a0, a0);
    u64 cross = mul_wide(a0, a1);                     // appears twice
    u64 p11 = mul_wide(a1, a1);
    // x = p00 + 2*cross*2^32 + p11*2^64
    u64 c2lo = cross << 1;                            // low 64 bits of 2*cross
    u32 c2hi = (u32)(cross >> 63);                    // bit 64
    u64 mid = c2lo + hi32(p00);
    u32 carry = mid < c2lo;
    u64 lo = pack64(lo32(p00), lo32(mid));
    u64 hi = p11 + hi32(mid) + ((u64)(c2hi + carry) << 32);
    return gl_reduce128(lo, hi);
}

GL_D u64 gl_pow7(u64 x) {
    u64 x2 = gl_sqr(x);
    u64 x4 = gl_sqr(x2);
    u64 x3 = gl_mul(x2, x);
    return gl_mul(x3, x4);
}

#ifndef __CUDACC_RTC__          // host-side helpers (tables, setup): not part of run-time compiled kernels
__host__ __device__ inline u64 gl_mul_slow(u64 a, u64 b) {     // host+device helper (tables, setup)
#ifdef __CUDA_ARCH__
    return gl_canon(gl_mul(a, b));
#else
    unsigned __int128 x = (unsigned __int128)a * b;
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hh = hi >> 32, hl = hi & GL_EPS;
    u64 t = lo - hh;
    if (lo < hh) t -= GL_EPS;
    u64 m = hl * GL_EPS;
    u64 r = t + m;
    if (r < t) r += GL_EPS;
    return r >= GL_P ? r - GL_P : r;
#endif
}

inline u64 gl_pow_host(u64 a, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = gl_mul_slow(r, a);
        a = gl_mul_slow(a, a);
        e >>= 1;
    }
    return r;
}
inline u64 gl_inv_host(u64 a) { return gl_pow_host(a, GL_P - 2); }
inline u64 gl_root_of_unity_host(unsigned log_n) {
    u64 r = GL_POWER_OF_TWO_GENERATOR;
    for (unsigned i = log_n; i < 32; i++) r = gl_mul_slow(r, r);
    return r;
}

#endif

GL_D u64 gl_pow(u64 a, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, a);
        a = gl_sqr(a);
        e >>= 1;
    }
    return r;
}
GL_D u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }

// ---- quadratic extension F[x]/(x^2 - 7) (plonky2_field extension/quadratic.rs; bound at
// contracts/lib/succinctx/plonky2x/core/src/backend/circuit/config.rs:41) --------------------------
struct gl2 {
    u64 a, b;       // a + b x
};
__host__ __device__ __forceinline__ gl2 gl2_make(u64 a, u64 b) { gl2 r; r.a = a; r.b = b; return r; }
GL_D gl2 gl2_add(gl2 x, gl2 y) { return gl2_make(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
GL_D gl2 gl2_sub(gl2 x, gl2 y) { return gl2_make(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
GL_D gl2 gl2_mul(gl2 x, gl2 y) {
    // (a0 b0 + 7 a1 b1) + (a0 b1 + a1 b0) x, two lazy dot products
    GlAcc t0, t1;
    gl_acc_init(t0, 0); gl_acc_init(t1, 0);
    u64 b7 = gl_mul_cc(y.b, 7);
    gl_acc_mad(t0, x.a, y.a); gl_acc_mad(t0, x.b, b7);
    gl_acc_mad(t1, x.a, y.b); gl_acc_mad(t1, x.b, y.a);
    return gl2_make(gl_acc_reduce(t0), gl_acc_reduce(t1));
}
GL_D gl2 gl2_mul_base(gl2 x, u64 k) { return gl2_make(gl_mul_cc(x.a, k), gl_mul_cc(x.b, k)); }
GL_D gl2 gl2_add_base(gl2 x, u64 k) { return gl2_make(gl_add(x.a, k), x.b); }
GL_D gl2 gl2_canon(gl2 x) { return gl2_make(gl_canon(x.a), gl_canon(x.b)); }
GL_D gl2 gl2_pow(gl2 x, u64 e) {
    gl2 r = gl2_make(1, 0);
    while (e) {
        if (e & 1) r = gl2_mul(r, x);
        x = gl2_mul(x, x);
        e >>= 1;
    }
    return r;
}
GL_D gl2 gl2_inv(gl2 x) {
    // 1/(a + b x) = (a - b x) / (a^2 - 7 b^2)
    u64 norm = gl_sub(gl_mul_cc(x.a, x.a), gl_mul_cc(7, gl_mul_cc(x.b, x.b)));
    u64 ni = gl_inv(norm);
    return gl2_make(gl_mul_cc(x.a, ni), gl_mul_cc(gl_neg(x.b), ni));
}
