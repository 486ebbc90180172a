// PoseidonBN128Hash on sm_100a: the wrapper-stage Merkle hasher.
//
// Replaces, for MerkleTree::<GoldilocksField, PoseidonBN128Hash>::new and the hasher surface it calls,
//   contracts/lib/succinctx/plonky2x/core/src/backend/wrapper/poseidon_bn128.rs:20-110   (permution)
//   contracts/lib/succinctx/plonky2x/core/src/backend/wrapper/plonky2_config.rs:128-197  (hash_no_pad, hash_or_noop, two_to_one)
// Digests are one BN254 scalar = 32 little-endian bytes = 4 u64 words (Fr::to_repr), canonical, so the trees use the same
// interleaved `digests` layout and the same path / query kernels as the Goldilocks-Poseidon trees (merkle.cu).
//
// Arithmetic: Montgomery form on 8 x 32-bit limbs; a product is 128 IMAD.WIDE.U32 (CIOS, one 32-bit reduction step per
// operand limb) and runs as a real call (noinline) so the rolled round loops stay inside the instruction cache.
// One thread per leaf / sibling pair; tables (C, S, M, P: 16 KB, Montgomery form) sit in __constant__ memory and are
// indexed warp-uniformly.
#include "bn128_tables.h"
#include "common.cuh"

namespace {

struct Fr { u32 v[8]; };

#define BN_N0 0xf0000001u
#define BN_N1 0x43e1f593u
#define BN_N2 0x79b97091u
#define BN_N3 0x2833e848u
#define BN_N4 0x8181585du
#define BN_N5 0xb85045b6u
#define BN_N6 0xe131a029u
#define BN_N7 0x30644e72u
#define BN_N0INV 0xefffffffu            // -N^-1 mod 2^32

__constant__ Fr c_bn_C[88];
__constant__ Fr c_bn_S[392];
__constant__ Fr c_bn_M[16];
__constant__ Fr c_bn_P[16];
__constant__ Fr c_bn_R2;                // 2^512 mod N: to Montgomery form

GL_D u32 bn_n(int j) {
    switch (j) {
        case 0: return BN_N0; case 1: return BN_N1; case 2: return BN_N2; case 3: return BN_N3;
        case 4: return BN_N4; case 5: return BN_N5; case 6: return BN_N6; default: return BN_N7;
    }
}

// r = t - N if t >= N else t   (t < 2N)
GL_D Fr fr_cond_sub(const u32 t[8]) {
    u32 d[8], borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(borrow)
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
          "n"(BN_N0), "n"(BN_N1), "n"(BN_N2), "n"(BN_N3), "n"(BN_N4), "n"(BN_N5), "n"(BN_N6), "n"(BN_N7));
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = borrow ? t[i] : d[i];
    return r;
}

GL_D Fr fr_add(const Fr& a, const Fr& b) {       // a, b < N < 2^254: the sum fits 8 limbs
    u32 t[8];
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    return fr_cond_sub(t);
}

// Montgomery product a b 2^-256 mod N, CIOS over 32-bit limbs.  Every step is x*y + t + carry <= 2^64 - 1: one IMAD.WIDE
// with the 32-bit t riding in its addend plus a two-word carry add.
__device__ __noinline__ Fr fr_mul(Fr a, Fr b) {
    u32 t[9];
#pragma unroll
    for (int j = 0; j < 9; j++) t[j] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u64 carry = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            u64 acc = mad_wide(a.v[j], b.v[i], (u64)t[j]) + carry;
            t[j] = lo32(acc);
            carry = hi32(acc);
        }
        u64 s = (u64)t[8] + carry;               // t < 2N < 2^255 after every outer step: no tenth limb
        t[8] = lo32(s);
        const u32 m = t[0] * BN_N0INV;
        u64 acc = mad_wide(m, BN_N0, (u64)t[0]);
        carry = hi32(acc);
#pragma unroll
        for (int j = 1; j < 8; j++) {
            acc = mad_wide(m, bn_n(j), (u64)t[j]) + carry;
            t[j - 1] = lo32(acc);
            carry = hi32(acc);
        }
        s = (u64)t[8] + carry;
        t[7] = lo32(s);
        t[8] = hi32(s);
    }
    return fr_cond_sub(t);
}

GL_D Fr fr_pow5(const Fr& x) {                    // exp5, poseidon_bn128.rs:33-41
    Fr x2 = fr_mul(x, x);
    Fr x4 = fr_mul(x2, x2);
    return fr_mul(x4, x);
}

GL_D Fr fr_zero() {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}

// canonical little-endian words -> Montgomery form.  The value must be < N (digests are; packed leaves are < 2^192).
GL_D Fr fr_from_words(u64 w0, u64 w1, u64 w2, u64 w3) {
    Fr r;
    r.v[0] = lo32(w0); r.v[1] = hi32(w0); r.v[2] = lo32(w1); r.v[3] = hi32(w1);
    r.v[4] = lo32(w2); r.v[5] = hi32(w2); r.v[6] = lo32(w3); r.v[7] = hi32(w3);
    return fr_mul(r, c_bn_R2);
}
GL_D void fr_to_words(const Fr& a, u64 out[4]) {
    Fr one = fr_zero();
    one.v[0] = 1;
    Fr r = fr_mul(a, one);
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = pack64(r.v[2 * i], r.v[2 * i + 1]);
}

// mix(), poseidon_bn128.rs:96-110: result[i] = sum_j m[j][i] * state[j]
GL_D void bn_mix(Fr s[4], const Fr* __restrict__ m) {
    Fr r[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        r[i] = fr_mul(m[i], s[0]);
#pragma unroll
        for (int j = 1; j < 4; j++) r[i] = fr_add(r[i], fr_mul(m[4 * j + i], s[j]));
    }
#pragma unroll
    for (int i = 0; i < 4; i++) s[i] = r[i];
}

// `permution`, poseidon_bn128.rs:20-25 (state in Montgomery form)
GL_D void bn_permute(Fr s[4]) {
#pragma unroll
    for (int i = 0; i < 4; i++) s[i] = fr_add(s[i], c_bn_C[i]);
    // full_rounds(first): three rounds with M, the fourth with P          (:48-69)
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 4; i++) s[i] = fr_add(fr_pow5(s[i]), c_bn_C[4 * (r + 1) + i]);
        bn_mix(s, r == 3 ? c_bn_P : c_bn_M);
    }
    // partial_rounds                                                     (:71-94)
#pragma unroll 1
    for (int r = 0; r < 56; r++) {
        s[0] = fr_add(fr_pow5(s[0]), c_bn_C[20 + r]);
        const Fr* S = c_bn_S + 7 * r;
        Fr n0 = fr_mul(S[0], s[0]);
#pragma unroll
        for (int j = 1; j < 4; j++) n0 = fr_add(n0, fr_mul(S[j], s[j]));
#pragma unroll
        for (int k = 1; k < 4; k++) s[k] = fr_add(s[k], fr_mul(s[0], S[3 + k]));
        s[0] = n0;
    }
    // full_rounds(last)                                                  (:48-69, first = false)
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            s[i] = fr_pow5(s[i]);
            if (r < 3) s[i] = fr_add(s[i], c_bn_C[76 + 4 * r + i]);
        }
        bn_mix(s, c_bn_M);
    }
}

GL_D uint64_t pair_pos(uint64_t q, uint32_t lvl) { return 2 * (q * (2ULL << lvl) + (1ULL << lvl) - 1); }

// hash_or_noop over Goldilocks elements read through `get(i)` (any representative; canonicalised here):
// three elements = 24 little-endian bytes per scalar, three scalars per permutation into state[1..4]
template <class Get>
GL_D void bn_hash_or_noop(uint32_t len, Get get, u64 out[4]) {
    if (len <= 3) {                                  // plonky2_config.rs:176-187: the digest IS the leaf's bytes
#pragma unroll
        for (int i = 0; i < 3; i++) out[i] = (uint32_t)i < len ? gl_canon(get(i)) : 0;
        out[3] = 0;
        return;
    }
    Fr s[4];
#pragma unroll
    for (int i = 0; i < 4; i++) s[i] = fr_zero();
    for (uint32_t off = 0; off < len; off += 9) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const uint32_t e = off + 3 * j;
            if (e < len) {                           // a short last chunk keeps the older lanes (overwrite mode)
                u64 w0 = gl_canon(get(e));
                u64 w1 = e + 1 < len ? gl_canon(get(e + 1)) : 0;
                u64 w2 = e + 2 < len ? gl_canon(get(e + 2)) : 0;
                s[j + 1] = fr_from_words(w0, w1, w2, 0);
            }
        }
        bn_permute(s);
    }
    fr_to_words(s[0], out);
}

GL_D void store4(u64* dst, const u64 w[4]) {
    reinterpret_cast<ulonglong2*>(dst)[0] = make_ulonglong2(w[0], w[1]);
    reinterpret_cast<ulonglong2*>(dst)[1] = make_ulonglong2(w[2], w[3]);
}

template <bool COL_MAJOR>
__global__ void __launch_bounds__(128) bn_leaf_kernel(const u64* __restrict__ leaves, uint64_t stride, uint64_t N, uint32_t c,
                                                      uint32_t sub_bits, u64* __restrict__ digests, u64* __restrict__ cap) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= N) return;
    const u64* src = COL_MAJOR ? leaves + row : leaves + row * c;
    const uint64_t step = COL_MAJOR ? stride : 1;
    u64 d[4];
    bn_hash_or_noop(c, [&](uint32_t i) { return src[(uint64_t)i * step]; }, d);
    u64* dst;
    if (sub_bits == 0) {
        dst = cap + 4 * row;
    } else {
        const uint64_t sub = 1ULL << sub_bits;
        const uint64_t sidx = row >> sub_bits, jj = row & (sub - 1);
        dst = digests + 4 * (sidx * (2 * sub - 2) + 4 * (jj >> 1) + (jj & 1));
    }
    store4(dst, d);
}

// two_to_one, plonky2_config.rs:189-196: permute([0, 0, left, right])[0]
__global__ void __launch_bounds__(128) bn_level_kernel(u64* __restrict__ digests, u64* __restrict__ cap, uint32_t lvl,
                                                       uint32_t sub_bits, uint64_t total_pairs) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_pairs) return;
    const uint32_t pair_bits = sub_bits - lvl - 1;
    const uint64_t sidx = t >> pair_bits, q = t & ((1ULL << pair_bits) - 1);
    const uint64_t sub = 1ULL << sub_bits;
    u64* blk = digests + 4 * sidx * (2 * sub - 2);
    const u64* pr = blk + 4 * pair_pos(q, lvl);
    Fr s[4];
    s[0] = fr_zero(); s[1] = fr_zero();
    s[2] = fr_from_words(pr[0], pr[1], pr[2], pr[3]);
    s[3] = fr_from_words(pr[4], pr[5], pr[6], pr[7]);
    bn_permute(s);
    u64 d[4];
    fr_to_words(s[0], d);
    u64* dst = (pair_bits == 0) ? cap + 4 * sidx : blk + 4 * (pair_pos(q >> 1, lvl + 1) + (q & 1));
    store4(dst, d);
}

__global__ void __launch_bounds__(128) bn_permute_kernel(const u64* __restrict__ in, uint64_t count, u64* __restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    Fr s[4];
#pragma unroll
    for (int i = 0; i < 4; i++) s[i] = fr_from_words(in[16 * t + 4 * i], in[16 * t + 4 * i + 1], in[16 * t + 4 * i + 2],
                                                     in[16 * t + 4 * i + 3]);
    bn_permute(s);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u64 d[4];
        fr_to_words(s[i], d);
        store4(out + 16 * t + 4 * i, d);
    }
}

__global__ void __launch_bounds__(128) bn_hash_kernel(const u64* __restrict__ in, uint64_t count, uint32_t len, bool or_noop,
                                                      u64* __restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const u64* src = in + t * len;
    u64 d[4];
    if (or_noop || len > 3) {
        bn_hash_or_noop(len, [&](uint32_t i) { return src[i]; }, d);
    } else {                                         // hash_no_pad of a short input still permutes (once; never for len 0)
        Fr s[4];
#pragma unroll
        for (int i = 0; i < 4; i++) s[i] = fr_zero();
        if (len) {
            s[1] = fr_from_words(gl_canon(src[0]), len > 1 ? gl_canon(src[1]) : 0, len > 2 ? gl_canon(src[2]) : 0, 0);
            bn_permute(s);
        }
        fr_to_words(s[0], d);
    }
    store4(out + 4 * t, d);
}

void split_words(const Bn128Fr& a, Fr* out) {
    for (int i = 0; i < 4; i++) { out->v[2 * i] = (u32)a.l[i]; out->v[2 * i + 1] = (u32)(a.l[i] >> 32); }
}
template <size_t K>
cudaError_t upload(const Fr (&sym)[K], const Bn128Fr* src) {
    static Fr tmp[K];
    for (size_t i = 0; i < K; i++) split_words(bn128_to_mont(src[i]), &tmp[i]);
    return cudaMemcpyToSymbol(sym, tmp, sizeof tmp);
}

}  // namespace

static const Bn128Tables* bn128_tables_host() {
    static Bn128Tables t;
    static bool ok = bn128_derive_tables(&t);
    return ok ? &t : nullptr;
}

int32_t bn128_module_init(vx_ctx* ctx) {
    const Bn128Tables* t = bn128_tables_host();
    VX_REQUIRE(t, "bn128: table derivation failed");
    VX_CUDA(upload(c_bn_C, t->C));
    VX_CUDA(upload(c_bn_S, t->S));
    VX_CUDA(upload(c_bn_M, t->M));
    VX_CUDA(upload(c_bn_P, t->P));
    Bn128Fr one = {{1, 0, 0, 0}};
    Bn128Fr r2 = bn128_to_mont(bn128_to_mont(one));      // 1 -> R -> R^2 (mod N)
    Fr r2w;
    split_words(r2, &r2w);
    VX_CUDA(cudaMemcpyToSymbol(c_bn_R2, &r2w, sizeof r2w));
    return VX_OK;
}

extern "C" int32_t vx_bn128_constants(uint64_t* c88, uint64_t* s392, uint64_t* m16, uint64_t* p16) {
    const Bn128Tables* t = bn128_tables_host();
    if (!t) { vx_set_error("vx_bn128_constants: table derivation failed"); return VX_EUNSUPPORTED; }
    if (c88) memcpy(c88, t->C, sizeof t->C);
    if (s392) memcpy(s392, t->S, sizeof t->S);
    if (m16) memcpy(m16, t->M, sizeof t->M);
    if (p16) memcpy(p16, t->P, sizeof t->P);
    return VX_OK;
}

int32_t bn128_merkle_build_device(vx_ctx* ctx, const u64* leaves, bool col_major, uint64_t stride, uint64_t N, uint32_t c,
                                  uint32_t cap_height, u64* digests, u64* cap, cudaEvent_t after_leaves) {
    const uint32_t log_N = ilog2(N);
    VX_REQUIRE((1ULL << log_N) == N, "merkle: leaf count %llu is not a power of two", (unsigned long long)N);
    VX_REQUIRE(cap_height <= log_N, "merkle: cap_height %u > log2(leaves) %u", cap_height, log_N);
    VX_REQUIRE(c >= 1, "merkle: empty leaves");
    const uint32_t sub_bits = log_N - cap_height;
    unsigned threads = 128;
    while (threads > 32 && (N + threads - 1) / threads < 4ULL * (uint64_t)ctx->sm_count) threads >>= 1;
    const unsigned blocks = (unsigned)((N + threads - 1) / threads);
    if (col_major) bn_leaf_kernel<true><<<blocks, threads, 0, ctx->stream>>>(leaves, stride, N, c, sub_bits, digests, cap);
    else bn_leaf_kernel<false><<<blocks, threads, 0, ctx->stream>>>(leaves, stride, N, c, sub_bits, digests, cap);
    VX_LAUNCH_COUNT(ctx, 1);
    if (after_leaves) VX_CUDA(cudaEventRecord(after_leaves, ctx->stream));
    for (uint32_t lvl = 0; lvl < sub_bits; lvl++) {
        const uint64_t total_pairs = N >> (lvl + 1);
        unsigned th = 128;
        while (th > 32 && (total_pairs + th - 1) / th < 4ULL * (uint64_t)ctx->sm_count) th >>= 1;
        bn_level_kernel<<<(unsigned)((total_pairs + th - 1) / th), th, 0, ctx->stream>>>(digests, cap, lvl, sub_bits, total_pairs);
        VX_LAUNCH_COUNT(ctx, 1);
    }
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

int32_t bn128_permute_device(vx_ctx* ctx, const u64* in, uint64_t count, u64* out) {
    bn_permute_kernel<<<(unsigned)((count + 63) / 64), 64, 0, ctx->stream>>>(in, count, out);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

int32_t bn128_hash_device(vx_ctx* ctx, const u64* in, uint64_t count, uint32_t len, bool or_noop, u64* out) {
    bn_hash_kernel<<<(unsigned)((count + 63) / 64), 64, 0, ctx->stream>>>(in, count, len, or_noop, out);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}
