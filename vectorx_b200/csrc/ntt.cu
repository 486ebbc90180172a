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
    return VX_OK;
}

void ntt_module_destroy(vx_ctx* ctx) {
    u64* ptrs[] = {ctx->w_lo, ctx->w_hi, ctx->wi_lo, ctx->wi_hi, ctx->g_lo, ctx->g_hi,
                   ctx->gi_lo, ctx->gi_hi, ctx->roots12, ctx->iroots12};
    for (u64* p : ptrs) if (p) cudaFree(p);
}

static TwiddleView tw_view(vx_ctx* ctx, bool inverse) {
    TwiddleView t;
    t.lo = inverse ? ctx->wi_lo : ctx->w_lo;
    t.hi = inverse ? ctx->wi_hi : ctx->w_hi;
    t.roots12 = inverse ? ctx->iroots12 : ctx->roots12;
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
    u64* base = data + ((uint64_t)blockIdx.y << log_n) + (blk << log_M) + i0_base;
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

int32_t ntt_dif_inplace(vx_ctx* ctx, u64* data, uint64_t count, uint32_t log_n, bool inverse) {
    if (log_n == 0 || count == 0) return VX_OK;
    VX_REQUIRE(log_n <= 32, "ntt: log_n %u exceeds the field's two-adicity", log_n);
    VX_REQUIRE(count < 65536, "ntt: too many transforms in one call (%llu)", (unsigned long long)count);
    TwiddleView tw = tw_view(ctx, inverse);
    uint32_t rem = log_n;
    while (rem > 12) {
        uint32_t A = rem - 8 < 8 ? rem - 8 : 8;
        uint64_t tiles = (1ULL << log_n) / ((1ULL << A) * NTT_TW);
        VX_REQUIRE(tiles < (1ULL << 31), "ntt: transform too large");
        dim3 grid((unsigned)tiles, (unsigned)count);
        size_t smem = (size_t)(1u << A) * NTT_PITCH * sizeof(u64);
        ntt_strided_pass<<<grid, 256, smem, ctx->stream>>>(data, log_n, rem, A, tw);
        VX_LAUNCH_COUNT(ctx, 1);
        rem -= A;
    }
    // a CTA owns 2^chunk_bits contiguous elements = one or more whole 2^rem blocks
    uint32_t chunk_bits = rem < 10 ? 10 : rem;
    uint64_t total = count << log_n;
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
