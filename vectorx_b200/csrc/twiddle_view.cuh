// Root-of-unity tables as the kernels see them, and the index bit reversal.  No host headers: also compiled by NVRTC.
#pragma once
#include "gl.cuh"

struct TwiddleView {     // passed by value to kernels
    const u64* lo;       // W^(E & 0xffff)
    const u64* hi;       // W^((E >> 16) << 16)
    const u64* roots12;  // w_4096^e, e < 2048
    const u64* full12;   // w_4096^e, e < 4096
};

__host__ __device__ static inline unsigned long long bitrev_u64(unsigned long long x, unsigned bits) {
#ifdef __CUDA_ARCH__
    return bits ? (__brevll(x) >> (64 - bits)) : 0;
#else
    unsigned long long r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
#endif
}

// W^E for a 32-bit exponent (W of order 2^32): w_M^e = W^(e << (32 - log M))
#ifdef __CUDACC__
GL_D u64 tw_pow_view(const TwiddleView& tw, u32 E) {
    u64 h = __ldg(tw.hi + (E >> 16));
    u32 l = E & 0xffffu;
    return l ? gl_mul_cc(h, __ldg(tw.lo + l)) : h;
}
#endif
