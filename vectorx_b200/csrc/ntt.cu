// Goldilocks NTT / iNTT / coset LDE on sm_100a.
//
// Replaces plonky2_field v0.2.0 fft.rs (`fft_classic`, `ifft_with_options`) and the LDE part of
// plonky2 fri/oracle.rs (`PolynomialBatch::lde_values` = lde + coset_fft(g)); reached in the
// reference from contracts/lib/succinctx/plonky2x/core/src/backend/circuit/build.rs:69-75 and
// .../frontend/hash/curta/stark.rs:123-126.
//
// Structure (B200-first, not plonky2's per-column radix-2 loop):
//  * every transform is an in-place decimation-in-frequency network: natural order in,
//    bit-reversed order out.  plonky2's Merkle leaf j is LDE point bitrev(j), so the LDE lands in
//    leaf order with no separate reverse_index_bits / transpose pass;
//  * a size-2^k transform is split four-step style into passes of <= 8 bits (strided tiles of
//    2^A x 16 elements: every global access is a full 128-byte line) and a final contiguous pass of
//    <= 12 bits; each pass runs its butterflies in shared memory;
//  * the LDE over the coset g<w_N> is 2^rate_bits independent size-n coset transforms of the same
//    coefficients: point i = 2^r k + rho is the n-point transform of (g w_N^rho)^m c_m at k, and its
//    leaf-order block is bitrev_r(rho).
#include "common.cuh"

// ------------------------------------------------------------------------------------------------ tables
static int32_t upload(vx_ctx* ctx, const std::vector<u64>& h, u64** d) {
    VX_CUDA(cudaMalloc(d, h.size() * sizeof(u64)));
    VX_CUDA(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

static void two_level(u64 base, unsigned lo_bits, unsigned hi_count, std::vector<u64>& lo, std::vector<u64>& hi) {
    lo.resize(1u << lo_bits);
    hi.resize(hi_count);
    u64 acc = 1;
    for (size_t i = 0; i < lo.size(); i++) { lo[i] = acc; acc = gl_mul_slow(acc, base); }
    u64 step = acc, a2 = 1;       // base^(2^lo_bits)
    for (size_t i = 0; i < hi.size(); i++) { hi[i] = a2; a2 = gl_mul_slow(a2, step); }
}

int32_t ntt_module_init(vx_ctx* ctx) {
    std::vector<u64> lo, hi;
    const u64 W = GL_POWER_OF_TWO_GENERATOR, Wi = gl_inv_host(W);
    two_level(W, 16, 1u << 16, lo, hi);
    VX_CHECK(upload(ctx, lo, &ctx->w_lo)); VX_CHECK(upload(ctx, hi, &ctx->w_hi));
    two_level(Wi, 16, 1u << 16, lo, hi);
    VX_CHECK(upload(ctx, lo, &ctx->wi_lo)); VX_CHECK(upload(ctx, hi, &ctx->wi_hi));
    const u64 G = GL_GENERATOR, Gi = gl_inv_host(G);
    two_level(G, 12, 1u << 14, lo, hi);          // exponents < 2^26
    VX_CHECK(upload(ctx, lo, &ctx->g_lo)); VX_CHECK(upload(ctx, hi, &ctx->g_hi));
    two_level(Gi, 12, 1u << 14, lo, hi);
    VX_CHECK(upload(ctx, lo, &ctx->gi_lo)); VX_CHECK(upload(ctx, hi, &ctx->gi_hi));
    std::vector<u64> r(2048);
    u64 w12 = gl_root_of_unity_host(12), acc = 1;
    for (int i = 0; i < 2048; i++) { r[i] = acc; acc = gl_mul_slow(acc, w12); }
    VX_CHECK(upload(ctx, r, &ctx->roots12));
    u64 w12i = gl_inv_host(w12); acc = 1;
    for (int i = 0; i < 2048; i++) { r[i] = acc; acc = gl_mul_slow(acc, w12i); }
    VX_CHECK(upload(ctx, r, &ctx->iroots12));
    r.resize(4096);
    acc = 1;
    for (int i = 0; i < 4096; i++) { r[i] = acc; acc = gl_mul_slow(acc, w12); }
    VX_CHECK(upload(ctx, r, &ctx->roots12f));
    acc = 1;
    for (int i = 0; i < 4096; i++) { r[i] = acc; acc = gl_mul_slow(acc, w12i); }
    VX_CHECK(upload(ctx, r, &ctx->iroots12f));
    return VX_OK;
}

void ntt_module_destroy(vx_ctx* ctx) {
    u64* ptrs[] = {ctx->w_lo, ctx->w_hi, ctx->wi_lo, ctx->wi_hi, ctx->g_lo, ctx->g_hi,
                   ctx->gi_lo, ctx->gi_hi, ctx->roots12, ctx->iroots12, ctx->roots12f, ctx->iroots12f};
    for (u64* p : ptrs) if (p) cudaFree(p);
}

static TwiddleView tw_view(vx_ctx* ctx, bool inverse) {
    TwiddleView t;
    t.lo = inverse ? ctx->wi_lo : ctx->w_lo;
    t.hi = inverse ? ctx->wi_hi : ctx->w_hi;
    t.roots12 = inverse ? ctx->iroots12 : ctx->roots12;
    t.full12 = inverse ? ctx->iroots12f : ctx->roots12f;
    return t;
}

// W^E for a 32-bit exponent (W of order 2^32): w_M^e = W^(e << (32 - log M))
GL_D u64 tw_pow(const TwiddleView& tw, u32 E) {
    u64 h = __ldg(tw.hi + (E >> 16));
    u32 l = E & 0xffffu;
    return l ? gl_mul(h, __ldg(tw.lo + l)) : h;
}

// ------------------------------------------------------------------------------------------------ passes
#define NTT_TW 16          // tile width along the contiguous dimension (16 x 8 B = one 128 B line)
#define NTT_PITCH 17       // padded shared-memory row pitch (in u64)

// Strided pass: A DIF stages over the high A bits of each 2^log_M block, then the four-step
// twiddle w_M^(i0 * k1).  grid.x = tiles, grid.y = transforms.  smem: 2^A x NTT_PITCH u64.
__global__ void __launch_bounds__(256) ntt_strided_pass(u64* __restrict__ data, uint32_t log_n, uint32_t log_M,
                                                        uint32_t A, TwiddleView tw) {
    extern __shared__ u64 sm[];
    const uint32_t low = log_M - A;
    const uint32_t tiles_per_blk = 1u << (low - 4);
    const uint64_t blk = blockIdx.x / tiles_per_blk;
    const uint32_t i0_base = (blockIdx.x % tiles_per_blk) * NTT_TW;
    u64* base = data + (((uint64_t)blockIdx.z * gridDim.y + blockIdx.y) << log_n) + (blk << log_M) + i0_base;
    const uint32_t R = 1u << A;
    const uint32_t elems = R * NTT_TW;

    for (uint32_t e = threadIdx.x; e < elems; e += blockDim.x) {
        uint32_t j1 = e >> 4, t = e & 15;
        sm[j1 * NTT_PITCH + t] = base[((uint64_t)j1 << low) + t];
    }
    __syncthreads();
    for (uint32_t s = 0; s < A; s++) {
        const uint32_t half_bits = A - 1 - s;
        const uint32_t half = 1u << half_bits;
        for (uint32_t b = threadIdx.x; b < elems / 2; b += blockDim.x) {
            uint32_t t = b & 15, bb = b >> 4;
            uint32_t j = bb & (half - 1);
            uint32_t i = ((bb >> half_bits) << (half_bits + 1)) + j;
            u64 u = sm[i * NTT_PITCH + t], v = sm[(i + half) * NTT_PITCH + t];
            u64 w = __ldg(tw.roots12 + (j << (11 - half_bits)));      // w_{2 half}^j
            sm[i * NTT_PITCH + t] = gl_add(u, v);
            u64 d = gl_sub(u, v);
            sm[(i + half) * NTT_PITCH + t] = j ? gl_mul(d, w) : d;
        }
        __syncthreads();
    }
    for (uint32_t e = threadIdx.x; e < elems; e += blockDim.x) {
        uint32_t j1 = e >> 4, t = e & 15;
        uint32_t k1 = __brev(j1) >> (32 - A);
        uint32_t i0 = i0_base + t;
        u64 x = sm[j1 * NTT_PITCH + t];
        uint32_t E = (i0 * k1) << (32 - log_M);
        if (E) x = gl_mul(x, tw_pow(tw, E));
        base[((uint64_t)j1 << low) + t] = gl_canon(x);
    }
}

// Final pass: F DIF stages on contiguous 2^F blocks; a CTA owns `chunk` (>= 2^F) contiguous elements.
__global__ void __launch_bounds__(256) ntt_final_pass(u64* __restrict__ data, uint32_t F, uint32_t chunk_bits,
                                                      TwiddleView tw) {
    extern __shared__ u64 sm[];
    const uint32_t chunk = 1u << chunk_bits;
    u64* base = data + ((uint64_t)blockIdx.x << chunk_bits);
    for (uint32_t e = threadIdx.x; e < chunk; e += blockDim.x) sm[e] = base[e];
    __syncthreads();
    for (uint32_t s = 0; s < F; s++) {
        const uint32_t half_bits = F - 1 - s;
        const uint32_t half = 1u << half_bits;
        for (uint32_t b = threadIdx.x; b < chunk / 2; b += blockDim.x) {
            uint32_t j = b & (half - 1);
            uint32_t i = ((b >> half_bits) << (half_bits + 1)) + j;
            u64 u = sm[i], v = sm[i + half];
            u64 w = __ldg(tw.roots12 + (j << (11 - half_bits)));
            sm[i] = gl_add(u, v);
            u64 d = gl_sub(u, v);
            sm[i + half] = j ? gl_mul(d, w) : d;
        }
        __syncthreads();
    }
    for (uint32_t e = threadIdx.x; e < chunk; e += blockDim.x) base[e] = gl_canon(sm[e]);
}

// ------------------------------------------------------------------------------------------------ radix-16 register passes
// In Goldilocks w_64 = 8, so every root of unity of order <= 64 is a power of two: a 2^R-point DFT (R <= 4) needs no
// real multiplication, only products by the compile-time constants 2^(96 j / half) below (ptxas folds the zero halves).
// A "group" is R DIF stages at once on 2^R elements held in registers:
//   y_j = sum_k x_k w_{2^R}^{jk}   (multiplication-free),   then   y_j *= w_M^{b j}   (M = 2^R S, b = index below S),
// with y_j left at position bitrev_R(j) -- exactly what R radix-2 DIF stages over blocks of size M would produce.
// The inverse transform is the same network on x_{(-k) mod 2^R} with the inverse twiddle tables.
template <int S_BITS>
GL_D u64 gl_mul_2exp(u64 x) {      // x * 2^S_BITS mod p, S_BITS < 96
    constexpr u64 C = S_BITS < 64 ? (1ULL << (S_BITS & 63))
                                  : (((1ULL << ((S_BITS - 64) & 31)) << 32) - (1ULL << ((S_BITS - 64) & 31)));   // 2^k * (2^32 - 1)
    return gl_mul_cc(x, C);
}

template <int R, int STAGE, int J>
GL_D u64 dft_twiddle(u64 d) {      // d * w_{2 half}^J with half = 2^(R - 1 - STAGE): w_{2 half} = 2^(96 / half)
    constexpr int half = 1 << (R - 1 - STAGE);
    constexpr int sh = (96 / half) * J;
    if constexpr (J == 0) return d;
    else return gl_mul_2exp<sh>(d);
}

template <int R, int STAGE, int B, int J>
GL_D void dft_stage_pair(u64* x) {
    constexpr int half = 1 << (R - 1 - STAGE);
    u64 u = x[B + J], v = x[B + J + half];
    x[B + J] = gl_add_cc(u, v);
    x[B + J + half] = dft_twiddle<R, STAGE, J>(gl_sub_cc(u, v));
}
template <int R, int STAGE, int B, int J>
GL_D void dft_stage_js(u64* x) {
    constexpr int half = 1 << (R - 1 - STAGE);
    if constexpr (J < half) {
        dft_stage_pair<R, STAGE, B, J>(x);
        dft_stage_js<R, STAGE, B, J + 1>(x);
    }
}
template <int R, int STAGE, int B>
GL_D void dft_stage_blocks(u64* x) {
    constexpr int half = 1 << (R - 1 - STAGE);
    if constexpr (B < (1 << R)) {
        dft_stage_js<R, STAGE, B, 0>(x);
        dft_stage_blocks<R, STAGE, B + 2 * half>(x);
    }
}
template <int R, int STAGE>
GL_D void dft_stages(u64* x) {
    if constexpr (STAGE < R) {
        dft_stage_blocks<R, STAGE, 0>(x);
        dft_stages<R, STAGE + 1>(x);
    }
}
// in place: x[q] <- y_{bitrev_R(q)}
template <int R>
GL_D void dft_dif(u64* x, bool inverse) {
    if (inverse) {
#pragma unroll
        for (int k = 1; k < (1 << R) / 2; k++) { u64 t = x[k]; x[k] = x[(1 << R) - k]; x[(1 << R) - k] = t; }
    }
    dft_stages<R, 0>(x);
}
__host__ __device__ constexpr int brev_small(int q, int bits) {
    int r = 0;
    for (int i = 0; i < bits; i++) r |= ((q >> i) & 1) << (bits - 1 - i);
    return r;
}

struct LdeScale {               // coset scaling fused into the first pass of the LDE: x_m *= (g w_N^rho)^m
    const u64 *g_lo, *g_hi;     // g^m = g_hi[m >> 12] * g_lo[m & 4095]
    TwiddleView fwd;            // forward W tables for w_N^(rho m)
    uint32_t log_N, rate_bits, blk_first, blk_count;
    u64 step[8][16];            // step[b][k] = (g w_N^rho_b)^(k * 16 * 2^low), b < blk_count <= 8
};

// Strided pass: the top 8 bits of every 2^log_M block, as two radix-16 groups.  A CTA owns a tile of 256 rows (row
// stride 2^low elements, low = log_M - 8) x 16 contiguous elements (one 128-byte line per row); thread (r, t) holds
// rows r + 16 k of column t for the first group and rows 16 r + k for the second (exchange through shared memory).
// grid.x = tiles per transform, grid.y = transforms.  SCALE: in = coefficients (one transform per column), out = LDE
// (blk_count transforms per column), log_M = log_n.
template <bool SCALE>
__global__ void __launch_bounds__(256) ntt_strided256_kernel(const u64* in, u64* out, uint32_t log_n, uint32_t log_M,
                                                             TwiddleView tw, bool inverse,
                                                             const __grid_constant__ LdeScale sc) {
    __shared__ u64 sm[256 * NTT_PITCH];
    const uint32_t low = log_M - 8;
    const uint32_t tiles_per_blk = 1u << (low - 4);
    const uint64_t blk = blockIdx.x / tiles_per_blk;
    const uint32_t i0 = (blockIdx.x % tiles_per_blk) * NTT_TW + (threadIdx.x & 15);
    const uint32_t t = threadIdx.x & 15, r = threadIdx.x >> 4;
    const uint64_t tr = (uint64_t)blockIdx.z * gridDim.y + blockIdx.y;       // transform index (grid.y x grid.z)
    uint64_t src_tr = tr, dst_tr = tr;
    uint32_t cb = 0;
    if (SCALE) { cb = (uint32_t)(tr % sc.blk_count); src_tr = tr / sc.blk_count; }
    const u64* src = in + (src_tr << log_n) + (blk << log_M) + i0;
    u64* dst = out + (dst_tr << log_n) + (blk << log_M) + i0;
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = src[(uint64_t)(r + 16 * k) << low];
    const uint32_t b1 = (r << low) + i0;                      // index below S1 = 16 * 2^low
    if (SCALE) {
        const uint32_t rho = sc.rate_bits ? (__brev(sc.blk_first + cb) >> (32 - sc.rate_bits)) : 0;
        u64 f = gl_mul_cc(__ldg(sc.g_hi + (b1 >> 12)), __ldg(sc.g_lo + (b1 & 4095)));
        const u32 E = (b1 * rho) << (32 - sc.log_N);
        if (E) f = gl_mul_cc(f, tw_pow_view(sc.fwd, E));
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = gl_mul_cc(x[k], k ? gl_mul_cc(f, sc.step[cb][k]) : f);
    }
    dft_dif<4>(x, inverse);
#pragma unroll
    for (int q = 1; q < 16; q++) {
        const u32 E = (b1 * (u32)brev_small(q, 4)) << (32 - log_M);
        if (E) x[q] = gl_mul_cc(x[q], tw_pow_view(tw, E));
    }
#pragma unroll
    for (int q = 0; q < 16; q++) sm[(r + 16 * q) * NTT_PITCH + t] = x[q];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = sm[(16 * r + k) * NTT_PITCH + t];
    dft_dif<4>(x, inverse);
#pragma unroll
    for (int q = 1; q < 16; q++) {
        const u32 E = (i0 * (u32)brev_small(q, 4)) << (32 - (log_M - 4));      // w_{M/16}^(i0 j)
        if (E) x[q] = gl_mul_cc(x[q], tw_pow_view(tw, E));
    }
#pragma unroll
    for (int q = 0; q < 16; q++) dst[(uint64_t)(16 * r + q) << low] = gl_canon(x[q]);
}

// one radix-2^R group over a 4096-element chunk held in (padded) shared memory; S = elements below the group
#define NTT_PAD(i) ((i) + ((i) >> 4))
template <int R>
GL_D void ntt_smem_group(u64* sm, uint32_t s_bits, const u64* __restrict__ full12, bool inverse) {
    const uint32_t S = 1u << s_bits;
    for (uint32_t u = threadIdx.x; u < (4096u >> R); u += 256) {
        const uint32_t b = u & (S - 1);
        const uint32_t base = ((u >> s_bits) << (s_bits + R)) + b;
        u64 x[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); k++) x[k] = sm[NTT_PAD(base + ((uint32_t)k << s_bits))];
        dft_dif<R>(x, inverse);
        if (s_bits) {
#pragma unroll
            for (int q = 1; q < (1 << R); q++) {
                const uint32_t e = (b * (uint32_t)brev_small(q, R)) << (12 - s_bits - R);
                if (e) x[q] = gl_mul_cc(x[q], __ldg(full12 + e));
            }
        }
#pragma unroll
        for (int q = 0; q < (1 << R); q++) sm[NTT_PAD(base + ((uint32_t)q << s_bits))] = x[q];
    }
}

// Final pass: F DIF stages on contiguous 2^F blocks (F <= 12); a CTA owns 4096 contiguous elements.
__global__ void __launch_bounds__(256) ntt_final4096_kernel(u64* __restrict__ data, uint32_t F, TwiddleView tw, bool inverse) {
    __shared__ u64 sm[4096 + 256];
    u64* base = data + ((uint64_t)blockIdx.x << 12);
    for (uint32_t e = threadIdx.x; e < 4096; e += 256) sm[NTT_PAD(e)] = base[e];
    __syncthreads();
    uint32_t rem = F;
    const uint32_t r1 = ((F - 1) & 3) + 1;                  // first group takes F mod 4 bits (or 4)
    rem -= r1;
    switch (r1) {
        case 1: ntt_smem_group<1>(sm, rem, tw.full12, inverse); break;
        case 2: ntt_smem_group<2>(sm, rem, tw.full12, inverse); break;
        case 3: ntt_smem_group<3>(sm, rem, tw.full12, inverse); break;
        default: ntt_smem_group<4>(sm, rem, tw.full12, inverse); break;
    }
    __syncthreads();
    while (rem) {
        rem -= 4;
        ntt_smem_group<4>(sm, rem, tw.full12, inverse);
        __syncthreads();
    }
    for (uint32_t e = threadIdx.x; e < 4096; e += 256) base[e] = gl_canon(sm[NTT_PAD(e)]);
}

int32_t ntt_dif_inplace(vx_ctx* ctx, u64* data, uint64_t count, uint32_t log_n, bool inverse) {
    if (log_n == 0 || count == 0) return VX_OK;
    VX_REQUIRE(log_n <= 32, "ntt: log_n %u exceeds the field's two-adicity", log_n);
    TwiddleView tw = tw_view(ctx, inverse);
    uint32_t rem = log_n;
    const bool fast = !ctx->ntt_legacy;
    while (rem > 12) {
        uint64_t tiles;
        // transforms are spread over grid.y x grid.z (count must factor as gy * gz with gy a power of two <= 32768)
        uint64_t gy = count, gz = 1;
        while (gy > 32768 && (gy & 1) == 0) { gy >>= 1; gz <<= 1; }
        VX_REQUIRE(gy <= 65535 && gz <= 65535, "ntt: cannot tile %llu transforms over the grid", (unsigned long long)count);
        if (fast && rem >= 16) {                       // radix-16 register pass over the top 8 bits
            tiles = (1ULL << log_n) >> 12;
            VX_REQUIRE(tiles < (1ULL << 31), "ntt: transform too large");
            dim3 grid((unsigned)tiles, (unsigned)gy, (unsigned)gz);
            ntt_strided256_kernel<false><<<grid, 256, 0, ctx->stream>>>(data, data, log_n, rem, tw, inverse, LdeScale{});
            VX_LAUNCH_COUNT(ctx, 1);
            rem -= 8;
            continue;
        }
        uint32_t A = rem - 8 < 8 ? rem - 8 : 8;
        tiles = (1ULL << log_n) / ((1ULL << A) * NTT_TW);
        VX_REQUIRE(tiles < (1ULL << 31), "ntt: transform too large");
        dim3 grid((unsigned)tiles, (unsigned)gy, (unsigned)gz);
        size_t smem = (size_t)(1u << A) * NTT_PITCH * sizeof(u64);
        ntt_strided_pass<<<grid, 256, smem, ctx->stream>>>(data, log_n, rem, A, tw);
        VX_LAUNCH_COUNT(ctx, 1);
        rem -= A;
    }
    uint64_t total = count << log_n;
    if (fast && rem >= 1 && (total & 4095) == 0) {
        VX_REQUIRE((total >> 12) < (1ULL << 31), "ntt: too many blocks");
        ntt_final4096_kernel<<<(unsigned)(total >> 12), 256, 0, ctx->stream>>>(data, rem, tw, inverse);
        VX_LAUNCH_COUNT(ctx, 1);
        VX_CUDA(cudaGetLastError());
        return VX_OK;
    }
    // a CTA owns 2^chunk_bits contiguous elements = one or more whole 2^rem blocks
    uint32_t chunk_bits = rem < 10 ? 10 : rem;
    uint32_t max_bits = log_n + (uint32_t)__builtin_ctzll(count);     // largest power of two dividing total
    if (chunk_bits > max_bits) chunk_bits = max_bits;
    uint64_t blocks = total >> chunk_bits;
    VX_REQUIRE(blocks < (1ULL << 31), "ntt: too many blocks");
    ntt_final_pass<<<(unsigned)blocks, 256, (size_t)sizeof(u64) << chunk_bits, ctx->stream>>>(data, rem, chunk_bits, tw);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

// ------------------------------------------------------------------------------------------------ glue kernels
// out[col][bitrev(p)] = in[col][p] * scale   (undo the DIF ordering; scale = n^-1 for the iNTT)
__global__ void bitrev_scale_kernel(const u64* __restrict__ in, u64* __restrict__ out, uint32_t log_n, u64 scale) {
    uint64_t n = 1ULL << log_n;
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const u64* src = in + ((uint64_t)blockIdx.y << log_n);
    u64* dst = out + ((uint64_t)blockIdx.y << log_n);
    u64 v = src[p];
    if (scale != 1) v = gl_mul(v, scale);
    dst[bitrev_u64(p, log_n)] = gl_canon(v);
}

int32_t intt_batch(vx_ctx* ctx, u64* work, u64* coeffs_out, uint32_t c, uint32_t log_n) {
    VX_CHECK(ntt_dif_inplace(ctx, work, c, log_n, true));
    uint64_t n = 1ULL << log_n;
    u64 ninv = gl_inv_host(n % GL_P);
    dim3 grid((unsigned)((n + 255) / 256), c);
    bitrev_scale_kernel<<<grid, 256, 0, ctx->stream>>>(work, coeffs_out, log_n, ninv);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

// lde[col][b * n + m] = coeffs[col][m] * g^m * w_N^(rho m),  rho = bitrev_r(blk_first + b), b < blk_count
__global__ void lde_scale_kernel(const u64* __restrict__ coeffs, u64* __restrict__ lde, uint32_t log_n,
                                 uint32_t rate_bits, uint32_t blk_first, uint32_t blk_count,
                                 const u64* __restrict__ g_lo, const u64* __restrict__ g_hi, TwiddleView tw) {
    uint64_t n = 1ULL << log_n;
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    uint32_t col = blockIdx.y;
    uint32_t log_N = log_n + rate_bits;
    u64 x = coeffs[((uint64_t)col << log_n) + m];
    u64 gm = gl_mul(__ldg(g_hi + (m >> 12)), __ldg(g_lo + (m & 4095)));
    x = gl_mul(x, gm);
    u64* dst = lde + (uint64_t)col * blk_count * n + m;
    for (uint32_t b = 0; b < blk_count; b++) {
        uint32_t rho = rate_bits ? (__brev(blk_first + b) >> (32 - rate_bits)) : 0;
        u32 E = ((u32)m * rho) << (32 - log_N);            // (rho m mod N) scaled to a W exponent
        u64 f = E ? gl_mul(x, tw_pow(tw, E)) : x;
        dst[(uint64_t)b << log_n] = gl_canon(f);
    }
}

int32_t lde_batch(vx_ctx* ctx, const u64* coeffs, u64* lde_out, uint32_t c, uint32_t log_n, uint32_t rate_bits,
                  uint32_t blk_first, uint32_t blk_count) {
    VX_REQUIRE(log_n + rate_bits <= 26, "lde: 2^%u points exceeds the coset table (2^26)", log_n + rate_bits);
    VX_REQUIRE(blk_count >= 1 && blk_first + blk_count <= (1u << rate_bits), "lde: coset block range out of bounds");
    uint64_t n = 1ULL << log_n;
    if (!ctx->ntt_legacy && log_n >= 16 && blk_count <= 8) {
        // first pass fused with the coset scaling: reads the coefficients once per coset (L2-resident), writes the LDE
        LdeScale sc;
        memset(&sc, 0, sizeof sc);
        sc.g_lo = ctx->g_lo; sc.g_hi = ctx->g_hi; sc.fwd = tw_view(ctx, false);
        sc.log_N = log_n + rate_bits; sc.rate_bits = rate_bits; sc.blk_first = blk_first; sc.blk_count = blk_count;
        const u64 wN = gl_root_of_unity_host(log_n + rate_bits);
        for (uint32_t b = 0; b < blk_count; b++) {
            uint32_t rho = (uint32_t)bitrev_u64(blk_first + b, rate_bits);
            u64 base = gl_mul_slow(GL_GENERATOR, gl_pow_host(wN, rho));          // g w_N^rho
            u64 st = gl_pow_host(base, n >> 4), acc = 1;                         // ^(16 * 2^low), low = log_n - 8
            for (int k = 0; k < 16; k++) { sc.step[b][k] = acc; acc = gl_mul_slow(acc, st); }
        }
        uint64_t gy = (uint64_t)c * blk_count, gz = 1;
        while (gy > 32768 && (gy & 1) == 0) { gy >>= 1; gz <<= 1; }
        VX_REQUIRE(gy <= 65535, "lde: cannot tile %u x %u transforms over the grid", c, blk_count);
        dim3 grid((unsigned)(n >> 12), (unsigned)gy, (unsigned)gz);
        ntt_strided256_kernel<true><<<grid, 256, 0, ctx->stream>>>(coeffs, lde_out, log_n, log_n, sc.fwd, false, sc);
        VX_LAUNCH_COUNT(ctx, 1);
        VX_CUDA(cudaGetLastError());
        // remaining stages: every 2^(log_n - 8) block is an independent transform
        return ntt_dif_inplace(ctx, lde_out, ((uint64_t)c * blk_count) << 8, log_n - 8, false);
    }
    dim3 grid((unsigned)((n + 255) / 256), c);
    lde_scale_kernel<<<grid, 256, 0, ctx->stream>>>(coeffs, lde_out, log_n, rate_bits, blk_first, blk_count,
                                                    ctx->g_lo, ctx->g_hi, tw_view(ctx, false));
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return ntt_dif_inplace(ctx, lde_out, (uint64_t)c * blk_count, log_n, false);
}

// x[m] *= shift^m (arbitrary shift; used by the generic vx_ntt entry point and FRI layers)
__global__ void coset_scale_kernel(u64* __restrict__ data, uint32_t log_n, u64 shift) {
    uint64_t n = 1ULL << log_n;
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    u64* p = data + ((uint64_t)blockIdx.y << log_n) + m;
    *p = gl_canon(gl_mul(*p, gl_pow(shift, m)));
}

int32_t ntt_natural(vx_ctx* ctx, const u64* in, u64* out, uint32_t c, uint32_t log_n, bool inverse,
                    uint64_t coset_shift) {
    uint64_t n = 1ULL << log_n;
    size_t bytes = (size_t)c * n * sizeof(u64);
    DevBuf tmp;
    VX_CHECK(tmp.alloc(bytes, ctx->stream));
    VX_CUDA(cudaMemcpyAsync(tmp.p, in, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    dim3 grid((unsigned)((n + 255) / 256), c);
    bool coset = coset_shift > 1;
    if (!inverse && coset) {
        coset_scale_kernel<<<grid, 256, 0, ctx->stream>>>(tmp.p, log_n, coset_shift);
        VX_LAUNCH_COUNT(ctx, 1);
    }
    VX_CHECK(ntt_dif_inplace(ctx, tmp.p, c, log_n, inverse));
    u64 scale = inverse ? gl_inv_host(n % GL_P) : 1;
    bitrev_scale_kernel<<<grid, 256, 0, ctx->stream>>>(tmp.p, out, log_n, scale);
    VX_LAUNCH_COUNT(ctx, 1);
    if (inverse && coset) {
        coset_scale_kernel<<<grid, 256, 0, ctx->stream>>>(out, log_n, gl_inv_host(coset_shift));
        VX_LAUNCH_COUNT(ctx, 1);
    }
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

int32_t coset_ntt_bitrev_inplace(vx_ctx* ctx, u64* data, uint32_t c, uint32_t log_n, uint64_t shift) {
    uint64_t n = 1ULL << log_n;
    if (shift > 1) {
        dim3 grid((unsigned)((n + 255) / 256), c);
        coset_scale_kernel<<<grid, 256, 0, ctx->stream>>>(data, log_n, shift);
        VX_LAUNCH_COUNT(ctx, 1);
        VX_CUDA(cudaGetLastError());
    }
    return ntt_dif_inplace(ctx, data, c, log_n, false);
}
