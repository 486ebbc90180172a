// Multiplication-free radix-2^R DFT kernels for the Goldilocks NTT, on lazily reduced 3-limb values.
//
// In Goldilocks 2^96 = -1, so 2 has order 192 and every root of unity of order <= 64 is a power of two (w_64 = 8,
// w_16 = 2^12): a 16-point DFT needs no field multiplication, only additions and products by 2^s.  Part of the NTT that
// replaces plonky2_field v0.2.0 fft.rs (`fft_classic`), reached in the reference from
// contracts/lib/succinctx/plonky2x/core/src/backend/circuit/build.rs:69-75.
//
// Representation inside a group: L3 = a + b 2^32 + c 2^64 with c a small SIGNED counter, i.e. a 96-bit two's-complement
// integer that is only congruent to the field element.  Additions and subtractions are three-instruction carry chains with
// no modular fix-up (a reduced add / sub pair costs 18 instructions); a product by 2^s is two IMAD.WIDE (FMA pipe) plus a
// fold through phi = 2^32: phi^2 = phi - 1, phi^3 = -1.  One reduction to a u64 representative happens per element at the
// end of a group, right before the general twiddle multiplication.
//
// The header compiles for the host too (plain C++ restatement of the same limb arithmetic, no PTX) so that
// tests/test_ntt_l3_host.py can check the fold formulas and the DFT network against a naive DFT on the CPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define L3_HD __host__ __device__ __forceinline__
#else
#define L3_HD inline
#endif

typedef unsigned int l3_u32;
typedef unsigned long long l3_u64;

struct L3 {
    l3_u32 a, b;
    int c;
};

L3_HD L3 l3_from(l3_u64 x) {
    L3 r;
    r.a = (l3_u32)x;
    r.b = (l3_u32)(x >> 32);
    r.c = 0;
    return r;
}

// host restatement helpers (the device path never touches them)
#if !defined(__CUDA_ARCH__)
inline __int128 l3_val(L3 x) { return (__int128)x.a + ((__int128)x.b << 32) + ((__int128)x.c * ((__int128)1 << 64)); }
inline L3 l3_pack(__int128 v) {
    L3 r;
    r.a = (l3_u32)(v & 0xffffffff);
    r.b = (l3_u32)((v >> 32) & 0xffffffff);
    r.c = (int)(long long)(v >> 64);
    return r;
}
#endif

L3_HD L3 l3_add(L3 x, L3 y) {
#if defined(__CUDA_ARCH__)
    L3 r;
    asm("add.cc.u32 %0, %3, %6;\n\taddc.cc.u32 %1, %4, %7;\n\taddc.u32 %2, %5, %8;"
        : "=r"(r.a), "=r"(r.b), "=r"(r.c)
        : "r"(x.a), "r"(x.b), "r"(x.c), "r"(y.a), "r"(y.b), "r"(y.c));
    return r;
#else
    return l3_pack(l3_val(x) + l3_val(y));
#endif
}
L3_HD L3 l3_sub(L3 x, L3 y) {
#if defined(__CUDA_ARCH__)
    L3 r;
    asm("sub.cc.u32 %0, %3, %6;\n\tsubc.cc.u32 %1, %4, %7;\n\tsubc.u32 %2, %5, %8;"
        : "=r"(r.a), "=r"(r.b), "=r"(r.c)
        : "r"(x.a), "r"(x.b), "r"(x.c), "r"(y.a), "r"(y.b), "r"(y.c));
    return r;
#else
    return l3_pack(l3_val(x) - l3_val(y));
#endif
}
// (w0 + w1 phi + cc phi^2) * phi^Q folded back to three limbs (phi^2 = phi - 1, phi^3 = -1); cc signed, |cc| < 2^30.
// The folded counter lands in [-2, 2].
template <int Q>
L3_HD L3 l3_fold(l3_u32 w0, l3_u32 w1, int cc) {
#if defined(__CUDA_ARCH__)
    L3 r;
    const int sx = cc >> 31;
    if (Q == 0) {           // (w0 - cc) + (w1 + cc) phi
        asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, 0, %6;\n\t"
            "add.cc.u32 %1, %1, %5;\n\taddc.u32 %2, %2, %6;"
            : "=&r"(r.a), "=&r"(r.b), "=&r"(r.c)
            : "r"(w0), "r"(w1), "r"(cc), "r"(sx));
    } else if (Q == 1) {    // -(w1 + cc) + (w0 + w1) phi
        asm("{\n\t.reg .u32 e0, e1, t1, t2, se;\n\t"
            "add.cc.u32 e0, %4, %5;\n\taddc.u32 e1, 0, %6;\n\t"
            "add.cc.u32 t1, %3, %4;\n\taddc.u32 t2, 0, 0;\n\t"
            "shr.s32 se, e1, 31;\n\t"
            "sub.cc.u32 %0, 0, e0;\n\tsubc.cc.u32 %1, t1, e1;\n\tsubc.u32 %2, t2, se;\n\t}"
            : "=r"(r.a), "=r"(r.b), "=r"(r.c)
            : "r"(w0), "r"(w1), "r"(cc), "r"(sx));
    } else {                // -(w0 + w1) + (w0 - cc) phi
        asm("{\n\t.reg .u32 t0, t1, u0, u1;\n\t"
            "add.cc.u32 t0, %3, %4;\n\taddc.u32 t1, 0, 0;\n\t"
            "sub.cc.u32 u0, %3, %5;\n\tsubc.u32 u1, 0, %6;\n\t"
            "sub.cc.u32 %0, 0, t0;\n\tsubc.cc.u32 %1, u0, t1;\n\tsubc.u32 %2, u1, 0;\n\t}"
            : "=r"(r.a), "=r"(r.b), "=r"(r.c)
            : "r"(w0), "r"(w1), "r"(cc), "r"(sx));
    }
    return r;
#else
    const __int128 PHI = (__int128)1 << 32;
    if (Q == 0) return l3_pack(((__int128)w0 - cc) + ((__int128)w1 + cc) * PHI);
    if (Q == 1) return l3_pack(-((__int128)w1 + cc) + ((__int128)w0 + w1) * PHI);
    return l3_pack(-((__int128)w0 + w1) + ((__int128)w0 - cc) * PHI);
#endif
}
// some u64 representative of x mod p (the host restatement returns the canonical one); |x.c| < 2^20
L3_HD l3_u64 l3_reduce(L3 x) {
#if defined(__CUDA_ARCH__)
    l3_u32 r0, r1;
    const int sx = x.c >> 31;
    // 96-bit (a, b, 0) - c + c 2^32; the sign words of c cancel in the top limb, which ends as w = carry - borrow in {-1, 0, 1}
    asm("{\n\t.reg .u32 t0, t1, w, s;\n\t"
        "sub.cc.u32 t0, %2, %4;\n\tsubc.cc.u32 t1, %3, %5;\n\tsubc.u32 w, 0, 0;\n\t"
        "add.cc.u32 t1, t1, %4;\n\taddc.u32 w, w, 0;\n\t"
        "shr.s32 s, w, 31;\n\t"
        "sub.cc.u32 %0, t0, w;\n\tsubc.u32 t1, t1, s;\n\tadd.u32 %1, t1, w;\n\t}"         // + w eps = (w << 32) - w
        : "=r"(r0), "=r"(r1)
        : "r"(x.a), "r"(x.b), "r"(x.c), "r"(sx));
    return ((l3_u64)r1 << 32) | r0;
#else
    const __int128 P = ((__int128)0xFFFFFFFF00000001ULL);
    __int128 v = l3_val(x) % P;
    if (v < 0) v += P;
    return (l3_u64)v;
#endif
}
// (a + b phi + c phi^2) 2^s = w0 + w1 phi + cc phi^2, 0 < s < 32: two IMAD.WIDE on the FMA pipe.  The passes are bound by
// the ALU pipe, and with a literal 2^s ptxas turns the products into funnel shifts (ALU); read from constant memory the
// factor stays an IMAD.WIDE operand.
#if defined(__CUDACC__)
static __constant__ l3_u32 c_l3_pow2[32] = {
    1u << 0,  1u << 1,  1u << 2,  1u << 3,  1u << 4,  1u << 5,  1u << 6,  1u << 7,  1u << 8,  1u << 9,  1u << 10,
    1u << 11, 1u << 12, 1u << 13, 1u << 14, 1u << 15, 1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 21,
    1u << 22, 1u << 23, 1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30, 1u << 31};
#endif
L3_HD void l3_shift(L3 x, int s, l3_u32& w0, l3_u32& w1, int& cc) {
#if defined(__CUDA_ARCH__)
    l3_u64 p0, p1;
    const l3_u32 k = c_l3_pow2[s];
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(x.a), "r"(k));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p1) : "r"(x.b), "r"(k), "l"(p0 >> 32));
#else
    const l3_u64 p0 = (l3_u64)x.a << s;
    const l3_u64 p1 = ((l3_u64)x.b << s) + (p0 >> 32);
#endif
    w0 = (l3_u32)p0;
    w1 = (l3_u32)p1;
#if defined(__CUDA_ARCH__)
    cc = (int)(l3_u32)(p1 >> 32) + x.c * (int)k;
#else
    cc = (int)(l3_u32)(p1 >> 32) + x.c * (1 << s);
#endif
}

// x * 2^SH, 0 <= SH < 96.  |x.c| must stay below 2^(30 - (SH & 31)) -- see the growth bound at l3_dft.
template <int SH>
L3_HD L3 l3_mul2exp(L3 x) {
    constexpr int s = SH & 31, q = SH >> 5;
    if (SH == 0) return x;
    l3_u32 w0, w1;
    int cc;
    if (s == 0) { w0 = x.a; w1 = x.b; cc = x.c; }
    else l3_shift(x, s, w0, w1, cc);
    return l3_fold<q>(w0, w1, cc);
}

// ---- 2^R-point DIF network, R <= 4: natural order in, x[q] <- y_{bitrev_R(q)} with y_j = sum_k x_k w^{jk}, w = 2^(192 / 2^R).
// Growth of the signed counter c: inputs have c = 0; an add / sub gives |c| <= |c_u| + |c_v| + 1 and every product by 2^s
// resets it to [-2, 2]: |c| <= 1, 5, 11, 23 after the four stages, and the shifted counters x.c 2^s (s <= 28, 24, 16 in
// stages 0, 1, 2) stay below 2^29.
template <int R, int STAGE, int B, int J>
L3_HD void l3_stage_pair(L3* x) {
    constexpr int half = 1 << (R - 1 - STAGE);
    const L3 u = x[B + J], v = x[B + J + half];
    x[B + J] = l3_add(u, v);
    x[B + J + half] = l3_mul2exp<(96 / half) * J>(l3_sub(u, v));
}
template <int R, int STAGE, int B, int J>
L3_HD void l3_stage_js(L3* x) {
    constexpr int half = 1 << (R - 1 - STAGE);
    if constexpr (J < half) {
        l3_stage_pair<R, STAGE, B, J>(x);
        l3_stage_js<R, STAGE, B, J + 1>(x);
    }
}
template <int R, int STAGE, int B>
L3_HD void l3_stage_blocks(L3* x) {
    constexpr int half = 1 << (R - 1 - STAGE);
    if constexpr (B < (1 << R)) {
        l3_stage_js<R, STAGE, B, 0>(x);
        l3_stage_blocks<R, STAGE, B + 2 * half>(x);
    }
}
template <int R, int STAGE>
L3_HD void l3_stages(L3* x) {
    if constexpr (STAGE < R) {
        l3_stage_blocks<R, STAGE, 0>(x);
        l3_stages<R, STAGE + 1>(x);
    }
}
// INV: the inverse transform is the same network on x_{(-k) mod 2^R}
template <int R, bool INV>
L3_HD void l3_dft(L3* x) {
    if (INV) {
#pragma unroll
        for (int k = 1; k < (1 << R) / 2; k++) { const L3 t = x[k]; x[k] = x[(1 << R) - k]; x[(1 << R) - k] = t; }
    }
    l3_stages<R, 0>(x);
}
