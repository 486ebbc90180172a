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
#include "ntt_l3.cuh"

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

static TwiddleView tw_view(vx_ctx* ctx, bool inverse);
__global__ void inner_table_kernel(u64* __restrict__ tab, TwiddleView tw);
static int32_t ntt_kernel_attributes();           // dynamic shared memory limits of the pipelined passes

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
    VX_CUDA(cudaMalloc(&ctx->inner_fwd, 256 * sizeof(u64)));
    VX_CUDA(cudaMalloc(&ctx->inner_inv, 256 * sizeof(u64)));
    inner_table_kernel<<<1, 256, 0, ctx->stream>>>(ctx->inner_fwd, tw_view(ctx, false));
    inner_table_kernel<<<1, 256, 0, ctx->stream>>>(ctx->inner_inv, tw_view(ctx, true));
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return ntt_kernel_attributes();
}

void ntt_module_destroy(vx_ctx* ctx) {
    u64* ptrs[] = {ctx->w_lo, ctx->w_hi, ctx->wi_lo, ctx->wi_hi, ctx->g_lo, ctx->g_hi,
                   ctx->gi_lo, ctx->gi_hi, ctx->roots12, ctx->iroots12, ctx->roots12f, ctx->iroots12f,
                   ctx->inner_fwd, ctx->inner_inv};
    for (u64* p : ptrs) if (p) cudaFree(p);
    for (auto& e : ctx->ntt_cache) cudaFree(e.p);
    ctx->ntt_cache.clear();
    ctx->ntt_cache_bytes = 0;
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

// ------------------------------------------------------------------------------------------------ derived tables
// Three kinds of device tables are derived once per shape and cached in the context (all canonical):
//   outer(log_M, inv)  [P < 256][i0 < M/256]  w_M^(+-i0 bitrev_8(P))      the four-step twiddles of a 256-row pass
//   scale(log_n, rate_bits, blk_first, blk_count)  [b][m < n]  (g w_N^rho_b)^m   the coset factors of the LDE
//   inner(inv)  [q < 16][r < 16]  w_256^(+-r bitrev_4(q))                 between the two radix-16 groups of a pass
// so that a twiddle costs one coalesced load and one multiplication (the two-level W tables cost two loads and two
// multiplications per twiddle).  Shapes whose table would be too big fall back to the two-level tables.
#define NTT_OUTER_MAX_LOG 24          // outer table: 8 * 2^log_M bytes (128 MB at 2^24, next to >= 8.6 GB of data)
#define NTT_SCALE_MAX_LOG 25          // scale table: 8 * blk_count * n bytes (256 MB at 2^25 entries)

__global__ void outer_table_kernel(u64* __restrict__ tab, uint32_t log_M, TwiddleView tw) {
    const uint32_t low = log_M - 8;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1ULL << log_M)) return;
    const uint32_t P = (uint32_t)(i >> low), i0 = (uint32_t)(i & ((1u << low) - 1));
    const uint32_t k1 = __brev(P) >> 24;
    tab[i] = gl_canon(tw_pow_view(tw, (i0 * k1) << (32 - log_M)));
}
__global__ void inner_table_kernel(u64* __restrict__ tab, TwiddleView tw) {
    const uint32_t q = threadIdx.x >> 4, r = threadIdx.x & 15;
    tab[threadIdx.x] = gl_canon(tw_pow_view(tw, (r * (__brev(q) >> 28)) << 24));      // w_256^e = W^(e << 24)
}
// tab[b][m] = g^m w_N^(rho_b m), rho_b = bitrev_r(blk_first + b) -- or rho_b = rho_explicit (one block: the folded cosets of
// lde_fold_kernel, whose exponent rho + 2^r kappa does not fit r bits)
__global__ void scale_table_kernel(u64* __restrict__ tab, uint32_t log_n, uint32_t rate_bits, uint32_t blk_first,
                                   const u64* __restrict__ g_lo, const u64* __restrict__ g_hi, TwiddleView tw,
                                   uint32_t rho_explicit, bool use_explicit) {
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >> log_n) return;
    const uint32_t b = blockIdx.y;
    const uint32_t rho = use_explicit ? rho_explicit : (rate_bits ? (__brev(blk_first + b) >> (32 - rate_bits)) : 0);
    u64 f = gl_mul_cc(__ldg(g_hi + (m >> 12)), __ldg(g_lo + (m & 4095)));
    const u32 E = ((u32)m * rho) << (32 - (log_n + rate_bits));
    if (E) f = gl_mul_cc(f, tw_pow_view(tw, E));
    tab[((uint64_t)b << log_n) + m] = gl_canon(f);
}

// Callers hold ctx->ntt_cache_mu.  A table is built on the context's main stream and the build is waited for, so any stream
// of the context may read it afterwards.
static const u64* cache_find(vx_ctx* lane, uint64_t key) {
    vx_ctx* ctx = lane->root;
    for (auto& e : ctx->ntt_cache) if (e.key == key) return e.p;
    return nullptr;
}
static int32_t cache_add(vx_ctx* lane, uint64_t key, size_t bytes, u64** out) {
    vx_ctx* ctx = lane->root;
    // derived tables are bounded by the number of distinct shapes a process commits; drop everything past 2 GB
    if (ctx->ntt_cache_bytes + bytes > (2ULL << 30)) {
        VX_CUDA(cudaDeviceSynchronize());
        for (auto& e : ctx->ntt_cache) cudaFree(e.p);
        ctx->ntt_cache.clear();
        ctx->ntt_cache_bytes = 0;
    }
    VX_CUDA(cudaMalloc(out, bytes));
    ctx->ntt_cache.push_back({key, *out, bytes});
    ctx->ntt_cache_bytes += bytes;
    return VX_OK;
}
static int32_t outer_table(vx_ctx* ctx, uint32_t log_M, bool inverse, const u64** out) {
    *out = nullptr;
    if (log_M > NTT_OUTER_MAX_LOG) return VX_OK;
    const uint64_t key = (1ULL << 60) | ((uint64_t)inverse << 8) | log_M;
    std::lock_guard<std::mutex> lk(ctx->root->ntt_cache_mu);
    if ((*out = cache_find(ctx, key))) return VX_OK;
    u64* p;
    VX_CHECK(cache_add(ctx, key, sizeof(u64) << log_M, &p));
    outer_table_kernel<<<(unsigned)(((1ULL << log_M) + 255) / 256), 256, 0, ctx->stream>>>(p, log_M, tw_view(ctx, inverse));
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = p;
    return VX_OK;
}
// rho_ext != UINT32_MAX: the one-block table of the folded coset with exponent rho_ext (see lde_fold_kernel)
static int32_t scale_table(vx_ctx* ctx, uint32_t log_n, uint32_t rate_bits, uint32_t blk_first, uint32_t blk_count,
                           const u64** out, uint32_t rho_ext = UINT32_MAX) {
    *out = nullptr;
    if (((uint64_t)blk_count << log_n) > (1ULL << NTT_SCALE_MAX_LOG)) return VX_OK;
    const bool ext = rho_ext != UINT32_MAX;
    const uint64_t key = ext ? ((3ULL << 60) | ((uint64_t)rho_ext << 16) | (rate_bits << 8) | log_n)
                             : ((2ULL << 60) | ((uint64_t)blk_count << 32) | ((uint64_t)blk_first << 16) | (rate_bits << 8) | log_n);
    std::lock_guard<std::mutex> lk(ctx->root->ntt_cache_mu);
    if ((*out = cache_find(ctx, key))) return VX_OK;
    u64* p;
    VX_CHECK(cache_add(ctx, key, ((size_t)blk_count << log_n) * sizeof(u64), &p));
    dim3 grid((unsigned)(((1ULL << log_n) + 255) / 256), blk_count);
    scale_table_kernel<<<grid, 256, 0, ctx->stream>>>(p, log_n, rate_bits, blk_first, ctx->g_lo, ctx->g_hi, tw_view(ctx, false),
                                                      ext ? rho_ext : 0u, ext);
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = p;
    return VX_OK;
}

// ------------------------------------------------------------------------------------------------ passes
#define NTT_TW 16          // tile width along the contiguous dimension (16 x 8 B = one 128 B line)
#define NTT_PITCH 17       // padded shared-memory row pitch (in u64)

// Final pass for shapes the 4096-element kernel does not cover (fewer than 4096 elements in total): F radix-2 DIF stages on
// contiguous 2^F blocks; a CTA owns `chunk` (>= 2^F) contiguous elements.
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

// ------------------------------------------------------------------------------------------------ radix-16 passes
// A "group" is 4 DIF stages at once on 16 register-resident elements: the multiplication-free 16-point DFT of ntt_l3.cuh
// (lazy 3-limb additions, products by powers of two), then ONE general twiddle per element.
__host__ __device__ constexpr int brev_small(int q, int bits) {
    int r = 0;
    for (int i = 0; i < bits; i++) r |= ((q >> i) & 1) << (bits - 1 - i);
    return r;
}

struct PassTables {
    const u64* inner;           // [16][16]
    const u64* outer;           // [256][M / 256] or nullptr (then the two-level W tables in `tw`)
    const u64* scale;           // SCALE: [blk_count][n] or nullptr (then g tables + `fwd`)
    const u64 *g_lo, *g_hi;
    TwiddleView tw, fwd;
    uint32_t log_N, rate_bits, blk_first, blk_count;
};

// Strided pass: the top 8 bits of every 2^log_M block as a 256-point DFT = two radix-16 groups with the w_256 twiddles
// between them, then the four-step twiddle w_M^(i0 k1).  A CTA owns a tile of 256 rows (row stride 2^low elements,
// low = log_M - 8) x 16 contiguous elements (one 128-byte line per row); thread (r, t) holds rows r + 16 k of column t for
// the first group and rows 16 r + k for the second (exchange through shared memory).  grid.x = tiles per transform,
// grid.y x grid.z = transforms.  SCALE: in = coefficients (one transform per column), out = LDE (blk_count transforms per
// column, each the coset transform of (g w_N^rho)^m c_m), log_M = log_n.
// DIRECT: the per-shape tables exist (tb.outer, and tb.scale when SCALE); otherwise the two-level W / g tables.
template <bool SCALE, bool INV, bool DIRECT>
__global__ void __launch_bounds__(256) ntt_pass256_kernel(const u64* in, u64* out, uint32_t log_n,
                                                          uint32_t log_M, const __grid_constant__ PassTables tb) {
    __shared__ u64 sm[256 * NTT_PITCH];
    const uint32_t low = log_M - 8;
    const uint32_t tiles_per_blk = 1u << (low - 4);
    const uint64_t blk = blockIdx.x / tiles_per_blk;
    const uint32_t t = threadIdx.x & 15, r = threadIdx.x >> 4;
    const uint32_t i0 = (blockIdx.x % tiles_per_blk) * NTT_TW + t;
    const uint64_t tr = (uint64_t)blockIdx.z * gridDim.y + blockIdx.y;       // transform index (grid.y x grid.z)
    uint64_t src_tr = tr;
    uint32_t cb = 0;
    if (SCALE) { cb = (uint32_t)(tr % tb.blk_count); src_tr = tr / tb.blk_count; }
    const size_t rs = (size_t)1 << low;                                        // row stride
    const u64* src = in + (src_tr << log_n) + (blk << log_M) + i0;
    u64* dst = out + (tr << log_n) + (blk << log_M) + i0;
    u64 v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = src[(r + 16 * k) * rs];
    if (SCALE) {
        if (DIRECT) {
            const u64* st = tb.scale + ((uint64_t)cb << log_n) + i0;
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = gl_mul_cc(v[k], __ldg(st + (r + 16 * k) * rs));
        } else {
            // shapes without a scale table: f = s^b1 from the two-level tables, then s^(b1 + k S1) = f * step^k
            const uint32_t b1 = (r << low) + i0;
            const uint32_t rho = tb.rate_bits ? (__brev(tb.blk_first + cb) >> (32 - tb.rate_bits)) : 0;
            u64 f = gl_mul_cc(__ldg(tb.g_hi + (b1 >> 12)), __ldg(tb.g_lo + (b1 & 4095)));
            const u32 E = (b1 * rho) << (32 - tb.log_N);
            if (E) f = gl_mul_cc(f, tw_pow_view(tb.fwd, E));
            const uint32_t S1 = 16u << low;
            u64 st = gl_mul_cc(__ldg(tb.g_hi + (S1 >> 12)), __ldg(tb.g_lo + (S1 & 4095)));
            const u32 Es = (S1 * rho) << (32 - tb.log_N);
            if (Es) st = gl_mul_cc(st, tw_pow_view(tb.fwd, Es));
#pragma unroll
            for (int k = 0; k < 16; k++) { v[k] = gl_mul_cc(v[k], f); f = gl_mul_cc(f, st); }
        }
    }
    L3 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = l3_from(v[k]);
    l3_dft<4, INV>(x);
    const u64* inner = tb.inner + r;
    sm[r * NTT_PITCH + t] = l3_reduce(x[0]);
#pragma unroll
    for (int q = 1; q < 16; q++) sm[(r + 16 * q) * NTT_PITCH + t] = gl_mul_cc(l3_reduce(x[q]), __ldg(inner + 16 * q));
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = l3_from(sm[(16 * r + k) * NTT_PITCH + t]);
    l3_dft<4, INV>(x);
    if (DIRECT) {
        const u64* ow = tb.outer + 16 * r * rs + i0;
        u64* d2 = dst + 16 * r * rs;
#pragma unroll
        for (int q = 0; q < 16; q++) d2[q * rs] = gl_mul_cc(l3_reduce(x[q]), __ldg(ow + q * rs));     // any representative: a final pass follows
    } else {
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const uint32_t k1 = (uint32_t)brev_small(q, 4) * 16 + (__brev(r) >> 28);      // bitrev_8(16 r + q)
            const u32 E = (i0 * k1) << (32 - log_M);
            u64 y = l3_reduce(x[q]);
            if (E) y = gl_mul_cc(y, tw_pow_view(tb.tw, E));
            dst[(16 * r + q) * rs] = y;
        }
    }
}

// one radix-2^R group over a 4096-element chunk; S = 2^s_bits elements below the group.  The chunk lives in (padded) shared
// memory; the FIRST group of a pass reads its inputs straight from global memory instead (all 16 loads of a thread in flight
// at once, every warp access a run of full 128-byte lines) and so needs no staging copy and no barrier before it.
#define NTT_PAD(i) ((i) + ((i) >> 4))
template <int R, bool INV, bool FROM_GLOBAL>
GL_D void ntt_smem_group(const u64* __restrict__ gin, u64* sm, uint32_t s_bits, const u64* __restrict__ full12) {
    constexpr int ITERS = 16 >> R;                          // (4096 >> R) units over 256 threads
    const uint32_t S = 1u << s_bits;
    u64 v[16];
    if (FROM_GLOBAL) {
#pragma unroll
        for (int it = 0; it < ITERS; it++) {
            const uint32_t u = threadIdx.x + 256 * it;
            const uint32_t base = ((u >> s_bits) << (s_bits + R)) + (u & (S - 1));
#pragma unroll
            for (int k = 0; k < (1 << R); k++) v[(it << R) + k] = gin[base + ((uint32_t)k << s_bits)];
        }
    }
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
        const uint32_t u = threadIdx.x + 256 * it;
        const uint32_t b = u & (S - 1);
        const uint32_t base = ((u >> s_bits) << (s_bits + R)) + b;
        L3 x[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); k++)
            x[k] = l3_from(FROM_GLOBAL ? v[(it << R) + k] : sm[NTT_PAD(base + ((uint32_t)k << s_bits))]);
        l3_dft<R, INV>(x);
        sm[NTT_PAD(base)] = l3_reduce(x[0]);
        if (s_bits) {
#pragma unroll
            for (int q = 1; q < (1 << R); q++) {
                const uint32_t e = (b * (uint32_t)brev_small(q, R)) << (12 - s_bits - R);      // w_(2^R S)^(b j), full12[0] = 1
                sm[NTT_PAD(base + ((uint32_t)q << s_bits))] = gl_mul_cc(l3_reduce(x[q]), __ldg(full12 + e));
            }
        } else {
#pragma unroll
            for (int q = 1; q < (1 << R); q++) sm[NTT_PAD(base + (uint32_t)q)] = l3_reduce(x[q]);
        }
    }
}

// Final pass: F DIF stages on contiguous 2^F blocks (F <= 12); a CTA owns 4096 contiguous elements.
#ifndef NTT_FINAL_MINB
#define NTT_FINAL_MINB 4
#endif
template <bool INV>
__global__ void __launch_bounds__(256, NTT_FINAL_MINB) ntt_final4096_kernel(u64* __restrict__ data, uint32_t F, const u64* __restrict__ full12) {
    __shared__ u64 sm[4096 + 256];
    u64* base = data + ((uint64_t)blockIdx.x << 12);
    uint32_t rem = F;
    const uint32_t r1 = ((F - 1) & 3) + 1;                  // first group takes F mod 4 bits (or 4)
    rem -= r1;
    switch (r1) {
        case 1: ntt_smem_group<1, INV, true>(base, sm, rem, full12); break;
        case 2: ntt_smem_group<2, INV, true>(base, sm, rem, full12); break;
        case 3: ntt_smem_group<3, INV, true>(base, sm, rem, full12); break;
        default: ntt_smem_group<4, INV, true>(base, sm, rem, full12); break;
    }
    __syncthreads();
    while (rem) {
        rem -= 4;
        ntt_smem_group<4, INV, false>(nullptr, sm, rem, full12);
        __syncthreads();
    }
    u64 o[16];
#pragma unroll
    for (int k = 0; k < 16; k++) o[k] = sm[NTT_PAD(threadIdx.x + 256 * k)];
#pragma unroll
    for (int k = 0; k < 16; k++) base[threadIdx.x + 256 * k] = gl_canon(o[k]);
}

// ---- the same pass as a persistent, double-buffered pipeline on the bulk-copy engine (TMA) ---------------------------------
// A CTA walks over chunks blockIdx.x, blockIdx.x + gridDim.x, ...  While it transforms chunk i, the 32 KB of chunk i + 1
// are already in flight (cp.async.bulk global -> shared, completion on an mbarrier) and the results of chunk i - 1 drain
// (cp.async.bulk shared -> global, bulk group): loads, stores and their address arithmetic leave the instruction stream and
// the DRAM latency that the one-shot kernel exposes once per CTA is hidden behind the butterflies.
GL_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
GL_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
GL_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
GL_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
GL_D void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
GL_D void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

#define NTT_TMA_SMEM (2 * 32768 + (4096 + 256) * 8)
template <bool INV>
__global__ void __launch_bounds__(256, 2) ntt_final4096_tma_kernel(u64* __restrict__ data, uint32_t chunks, uint32_t F,
                                                                   const u64* __restrict__ full12) {
    extern __shared__ __align__(128) unsigned char ntt_smraw[];
    u64* tile0 = reinterpret_cast<u64*>(ntt_smraw);                 // linear 4096-element tiles: load target, then store source
    u64* tile1 = reinterpret_cast<u64*>(ntt_smraw + 32768);
    u64* sm = reinterpret_cast<u64*>(ntt_smraw + 65536);            // padded working copy
    __shared__ uint64_t mbar[2];
    if (threadIdx.x == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t r1 = ((F - 1) & 3) + 1;                  // first group takes F mod 4 bits (or 4)
    uint32_t chunk = blockIdx.x;
    if (threadIdx.x == 0 && chunk < chunks) {
        mbar_expect_tx(&mbar[0], 32768);
        bulk_load(tile0, data + ((uint64_t)chunk << 12), 32768, &mbar[0]);
    }
    for (uint32_t it = 0; chunk < chunks; chunk += gridDim.x, it++) {
        u64* tile = (it & 1) ? tile1 : tile0;
        u64* other = (it & 1) ? tile0 : tile1;
        const uint32_t next = chunk + gridDim.x;
        if (threadIdx.x == 0 && next < chunks) {
            // `other` is the source of the store issued one iteration ago: it may be overwritten once that store has read it
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_expect_tx(&mbar[(it & 1) ^ 1], 32768);
            bulk_load(other, data + ((uint64_t)next << 12), 32768, &mbar[(it & 1) ^ 1]);
        }
        mbar_wait(&mbar[it & 1], (it >> 1) & 1);
        uint32_t rem = F - r1;
        switch (r1) {
            case 1: ntt_smem_group<1, INV, true>(tile, sm, rem, full12); break;
            case 2: ntt_smem_group<2, INV, true>(tile, sm, rem, full12); break;
            case 3: ntt_smem_group<3, INV, true>(tile, sm, rem, full12); break;
            default: ntt_smem_group<4, INV, true>(tile, sm, rem, full12); break;
        }
        __syncthreads();
        while (rem) {
            rem -= 4;
            ntt_smem_group<4, INV, false>(nullptr, sm, rem, full12);
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < 16; k++) tile[threadIdx.x + 256 * k] = gl_canon(sm[NTT_PAD(threadIdx.x + 256 * k)]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the copy engine
        __syncthreads();
        if (threadIdx.x == 0) bulk_store(data + ((uint64_t)chunk << 12), tile, 32768);
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static int32_t ntt_kernel_attributes() {
    VX_CUDA(cudaFuncSetAttribute(ntt_final4096_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_TMA_SMEM));
    VX_CUDA(cudaFuncSetAttribute(ntt_final4096_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_TMA_SMEM));
    return VX_OK;
}

// transforms are spread over grid.y x grid.z (count must factor as gy * gz with gy <= 65535)
static int32_t grid_yz(uint64_t count, uint64_t* gy, uint64_t* gz) {
    *gy = count; *gz = 1;
    while (*gy > 32768 && (*gy & 1) == 0) { *gy >>= 1; *gz <<= 1; }
    VX_REQUIRE(*gy <= 65535 && *gz <= 65535, "ntt: cannot tile %llu transforms over the grid", (unsigned long long)count);
    return VX_OK;
}

int32_t ntt_dif_inplace(vx_ctx* ctx, u64* data, uint64_t count, uint32_t log_n, bool inverse) {
    if (log_n == 0 || count == 0) return VX_OK;
    VX_REQUIRE(log_n <= 32, "ntt: log_n %u exceeds the field's two-adicity", log_n);
    TwiddleView tw = tw_view(ctx, inverse);
    uint32_t rem = log_n;
    while (rem > 12) {                                  // radix-16 register pass over the top 8 bits (rem - 8 >= 5 is left)
        uint64_t gy, gz;
        VX_CHECK(grid_yz(count, &gy, &gz));
        const uint64_t tiles = (1ULL << log_n) >> 12;
        VX_REQUIRE(tiles < (1ULL << 31), "ntt: transform too large");
        PassTables tb;
        memset(&tb, 0, sizeof tb);
        tb.inner = inverse ? ctx->inner_inv : ctx->inner_fwd;
        tb.tw = tw;
        VX_CHECK(outer_table(ctx, rem, inverse, &tb.outer));
        dim3 grid((unsigned)tiles, (unsigned)gy, (unsigned)gz);
        if (tb.outer) {
            if (inverse) ntt_pass256_kernel<false, true, true><<<grid, 256, 0, ctx->stream>>>(data, data, log_n, rem, tb);
            else ntt_pass256_kernel<false, false, true><<<grid, 256, 0, ctx->stream>>>(data, data, log_n, rem, tb);
        } else {
            if (inverse) ntt_pass256_kernel<false, true, false><<<grid, 256, 0, ctx->stream>>>(data, data, log_n, rem, tb);
            else ntt_pass256_kernel<false, false, false><<<grid, 256, 0, ctx->stream>>>(data, data, log_n, rem, tb);
        }
        VX_LAUNCH_COUNT(ctx, 1);
        rem -= 8;
    }
    uint64_t total = count << log_n;
    if ((total & 4095) == 0) {
        VX_REQUIRE((total >> 12) < (1ULL << 31), "ntt: too many blocks");
        const uint64_t chunks = total >> 12;
        // Measured (profiles/r02_ntt_ab.md): the pipelined form is CORRECT but slower here -- 0.52 against 0.43 ms for the
        // final pass of the config-1 LDE.  The pass is bound by the ALU pipe (83 % of its issue bound), not by memory
        // latency, and two 64 KB-staged CTAs per SM (16 warps) feed that pipe worse than four one-shot CTAs (32 warps).
        // It stays in the tree behind this switch (make EXTRA=-DNTT_FINAL_TMA=1) for shapes / parts where HBM is the bound.
#ifndef NTT_FINAL_TMA
#define NTT_FINAL_TMA 0
#endif
        if (NTT_FINAL_TMA && chunks >= 4ULL * (uint64_t)ctx->sm_count) {
            // enough chunks to keep a two-CTA-per-SM persistent grid busy: the pipelined form
            const unsigned grid = 2u * (unsigned)ctx->sm_count;
            if (inverse) ntt_final4096_tma_kernel<true><<<grid, 256, NTT_TMA_SMEM, ctx->stream>>>(data, (uint32_t)chunks, rem, tw.full12);
            else ntt_final4096_tma_kernel<false><<<grid, 256, NTT_TMA_SMEM, ctx->stream>>>(data, (uint32_t)chunks, rem, tw.full12);
        } else if (inverse) ntt_final4096_kernel<true><<<(unsigned)chunks, 256, 0, ctx->stream>>>(data, rem, tw.full12);
        else ntt_final4096_kernel<false><<<(unsigned)chunks, 256, 0, ctx->stream>>>(data, rem, tw.full12);
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

// The same reorder through 32 x 32 shared-memory tiles (log_n >= 10): an index is (hi | mid | lo) with 5-bit hi and lo, its
// reversal (rev lo | rev mid | rev hi).  A tile holds all (hi, lo) of one mid: rows are read with lo contiguous and written
// with rev(hi) contiguous, so both sides move 256-byte runs.  The element-wise scatter above writes one 8-byte word per
// 32-byte sector and, past a few MB per column, per TLB page: 59 ms of a 2^24 x 64 iNTT went there.
__global__ void __launch_bounds__(256) bitrev_scale_tiled_kernel(const u64* __restrict__ in, u64* __restrict__ out,
                                                                 uint32_t log_n, u64 scale) {
    __shared__ u64 tile[32][33];
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const uint32_t mid_bits = log_n - 10;
    const uint32_t mid = blockIdx.x;
    const uint32_t rmid = mid_bits ? (__brev(mid) >> (32 - mid_bits)) : 0;
    const u64* src = in + ((uint64_t)blockIdx.y << log_n);
    u64* dst = out + ((uint64_t)blockIdx.y << log_n);
#pragma unroll
    for (uint32_t hi = ty; hi < 32; hi += 8) tile[hi][tx] = src[((uint64_t)hi << (log_n - 5)) | ((uint64_t)mid << 5) | tx];
    __syncthreads();
    const uint32_t rb = __brev(tx) >> 27;                                  // hi = rev5(b), b = tx
#pragma unroll
    for (uint32_t a = ty; a < 32; a += 8) {
        u64 v = tile[rb][__brev(a) >> 27];                                 // lo = rev5(a)
        if (scale != 1) v = gl_mul(v, scale);
        dst[((uint64_t)a << (log_n - 5)) | ((uint64_t)rmid << 5) | tx] = gl_canon(v);
    }
}
static void launch_bitrev_scale(vx_ctx* ctx, const u64* in, u64* out, uint32_t c, uint32_t log_n, u64 scale) {
    const uint64_t n = 1ULL << log_n;
    if (log_n >= 10 && (n >> 10) < (1ULL << 31) && c <= 65535) {
        dim3 grid((unsigned)(n >> 10), c);
        bitrev_scale_tiled_kernel<<<grid, 256, 0, ctx->stream>>>(in, out, log_n, scale);
    } else {
        dim3 grid((unsigned)((n + 255) / 256), c);
        bitrev_scale_kernel<<<grid, 256, 0, ctx->stream>>>(in, out, log_n, scale);
    }
    VX_LAUNCH_COUNT(ctx, 1);
}

int32_t intt_batch(vx_ctx* ctx, u64* work, u64* coeffs_out, uint32_t c, uint32_t log_n) {
    VX_CHECK(ntt_dif_inplace(ctx, work, c, log_n, true));
    uint64_t n = 1ULL << log_n;
    u64 ninv = gl_inv_host(n % GL_P);
    launch_bitrev_scale(ctx, work, coeffs_out, c, log_n, ninv);
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

// Part q of the 2^kb equal parts of a coset's leaf block.  The leaves of a coset are the DIF (bit-reversed) output of the
// n-point transform of x_m = c_m s^m, s = g w_N^rho: part q holds the outputs k = 2^kb k' + kappa, kappa = bitrev_kb(q), and
//   X[2^kb k' + kappa] = sum_{m' < n'} ( sum_{j < 2^kb} x_{m' + j n'} w_n^{(m' + j n') kappa} ) w_{n'}^{m' k'},   n' = n / 2^kb,
// i.e. the n'-point transform of the FOLDED sequence y_{m'} = sum_j c_{m' + j n'} t^{m' + j n'} with t = g w_N^{rho + 2^r kappa}
// -- the evaluation of the polynomial on the smaller coset t <w_{n'}>.  A shard computes only its own part: 1 / 2^kb of the
// transform work for one extra read of the coefficients.
// `scale` (t^m, m < n; one coalesced load and one multiplication per coefficient) or, for shapes without a table, the
// two-level g / W tables.
__global__ void lde_fold_kernel(const u64* __restrict__ coeffs, u64* __restrict__ out, uint32_t log_n, uint32_t kb,
                                uint32_t rho_ext, uint32_t log_N, const u64* __restrict__ scale,
                                const u64* __restrict__ g_lo, const u64* __restrict__ g_hi, TwiddleView tw) {
    const uint32_t log_np = log_n - kb;
    const uint64_t mp = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (mp >> log_np) return;
    const u64* src = coeffs + ((uint64_t)blockIdx.y << log_n);
    GlAcc acc;
    gl_acc_init(acc, 0);
    for (uint32_t j = 0; j < (1u << kb); j++) {
        const uint32_t m = (uint32_t)mp + (j << log_np);
        u64 f;
        if (scale) {
            f = __ldg(scale + m);
        } else {
            f = gl_mul_cc(__ldg(g_hi + (m >> 12)), __ldg(g_lo + (m & 4095)));
            const u32 E = (m * rho_ext) << (32 - log_N);
            if (E) f = gl_mul_cc(f, tw_pow_view(tw, E));
        }
        gl_acc_mad(acc, src[m], f);                      // lazy dot product: one reduction per output
    }
    out[((uint64_t)blockIdx.y << log_np) + mp] = gl_canon(gl_acc_reduce(acc));
}

int32_t lde_batch(vx_ctx* ctx, const u64* coeffs, u64* lde_out, uint32_t c, uint32_t log_n, uint32_t rate_bits,
                  uint32_t blk_first, uint32_t blk_count, uint32_t fold_bits, uint32_t fold_index) {
    VX_REQUIRE(log_n <= 26 && log_n + rate_bits <= 32, "lde: 2^%u rows at rate_bits %u exceed the coset tables", log_n, rate_bits);
    VX_REQUIRE(blk_count >= 1 && blk_first + blk_count <= (1u << rate_bits), "lde: coset block range out of bounds");
    uint64_t n = 1ULL << log_n;
    if (fold_bits) {
        VX_REQUIRE(blk_count == 1 && fold_bits <= log_n && fold_index < (1u << fold_bits), "lde: bad sub-block %u of 2^%u", fold_index, fold_bits);
        const uint32_t rho = rate_bits ? (uint32_t)bitrev_u64(blk_first, rate_bits) : 0;
        const uint32_t kappa = (uint32_t)bitrev_u64(fold_index, fold_bits);
        const uint32_t log_np = log_n - fold_bits;
        const uint32_t rho_ext = rho + (kappa << rate_bits);
        const u64* scale = nullptr;
        VX_CHECK(scale_table(ctx, log_n, rate_bits, 0, 1, &scale, rho_ext));
        dim3 grid((unsigned)(((1ULL << log_np) + 255) / 256), c);
        lde_fold_kernel<<<grid, 256, 0, ctx->stream>>>(coeffs, lde_out, log_n, fold_bits, rho_ext, log_n + rate_bits, scale,
                                                       ctx->g_lo, ctx->g_hi, tw_view(ctx, false));
        VX_LAUNCH_COUNT(ctx, 1);
        VX_CUDA(cudaGetLastError());
        return ntt_dif_inplace(ctx, lde_out, c, log_np, false);
    }
    if (log_n >= 13) {
        // first pass fused with the coset scaling: reads the coefficients once per coset (L2-resident), writes the LDE
        PassTables tb;
        memset(&tb, 0, sizeof tb);
        tb.inner = ctx->inner_fwd;
        tb.tw = tb.fwd = tw_view(ctx, false);
        tb.g_lo = ctx->g_lo; tb.g_hi = ctx->g_hi;
        tb.log_N = log_n + rate_bits; tb.rate_bits = rate_bits; tb.blk_first = blk_first; tb.blk_count = blk_count;
        VX_CHECK(outer_table(ctx, log_n, false, &tb.outer));
        VX_CHECK(scale_table(ctx, log_n, rate_bits, blk_first, blk_count, &tb.scale));
        uint64_t gy, gz;
        VX_CHECK(grid_yz((uint64_t)c * blk_count, &gy, &gz));
        dim3 grid((unsigned)(n >> 12), (unsigned)gy, (unsigned)gz);
        if (tb.outer && tb.scale) ntt_pass256_kernel<true, false, true><<<grid, 256, 0, ctx->stream>>>(coeffs, lde_out, log_n, log_n, tb);
        else ntt_pass256_kernel<true, false, false><<<grid, 256, 0, ctx->stream>>>(coeffs, lde_out, log_n, log_n, tb);
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
    launch_bitrev_scale(ctx, tmp.p, out, c, log_n, scale);
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
