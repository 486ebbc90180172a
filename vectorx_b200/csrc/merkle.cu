// Poseidon Merkle tree on sm_100a.
//
// Replaces plonky2 v0.2.0 hash/merkle_tree.rs `MerkleTree::new` / `fill_digests_buf` /
// `fill_subtree`, hash/hashing.rs `hash_n_to_hash_no_pad` / `compress`, and the Hasher surface
// shown in-tree at contracts/lib/succinctx/plonky2x/core/src/backend/wrapper/plonky2_config.rs:130-196
// (hash_or_noop, two_to_one).  API shape exercised in the reference at
// .../backend/wrapper/poseidon_bn128.rs:217-220.
//
// Layout: `digests` is plonky2's interleaved layout -- per cap subtree a block of 2*sub-2 digests
// where the sibling pair q of layer l (0 = leaf digests) sits at 2*(q*2^(l+1) + 2^l - 1) + {0,1};
// roots go to `cap` only.  Kernels write straight into that layout so the tree a Rust caller
// downloads is byte-identical to MerkleTree { leaves, digests, cap }.
//
// Leaf hashing is one thread per leaf row.  With the LDE kept column-major in leaf order the
// per-column loads of a warp are 256 contiguous bytes (fully coalesced) with no transpose pass.
#include "common.cuh"
#include "poseidon.cuh"

// ------------------------------------------------------------------------------------------------
// Round constants: ChaCha8Rng::seed_from_u64(0), 360 x gen_range(0..p) (rand 0.8 widening-multiply
// rejection).  Derived here, independently of the test oracle, and audited by tests against it.
static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

static void chacha8_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                       key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                       (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t x[16];
    memcpy(x, in, sizeof x);
    auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
    };
    for (int dr = 0; dr < 4; dr++) {      // 8 rounds = 4 double rounds
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

void poseidon_round_constants_host(u64 out[360]) {
    uint32_t key[8];
    uint64_t pcg = 0;                     // rand_core::SeedableRng::seed_from_u64 (PCG32 expansion)
    for (int i = 0; i < 8; i++) {
        pcg = pcg * 6364136223846793005ULL + 11634580027462260723ULL;
        uint32_t xorshifted = (uint32_t)(((pcg >> 18) ^ pcg) >> 27);
        uint32_t rot = (uint32_t)(pcg >> 59);
        key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    }
    uint32_t words[16];
    int have = 0, used = 0;
    uint64_t counter = 0;
    auto next_u32 = [&]() {
        if (used == have) { chacha8_block(key, counter++, words); have = 16; used = 0; }
        return words[used++];
    };
    int n = 0;
    while (n < 360) {
        uint64_t lo = next_u32();
        uint64_t hi = next_u32();
        uint64_t v = lo | (hi << 32);
        unsigned __int128 m = (unsigned __int128)v * GL_P;
        if ((uint64_t)m <= GL_P - 1) out[n++] = (uint64_t)(m >> 64);
    }
}

static __device__ u64 g_pos_rc[31 * 12];      // global-memory copy of c_pos.rc for kernels that index it per lane

int32_t poseidon_module_init(vx_ctx* ctx) {
    u64 rc[360];
    poseidon_round_constants_host(rc);
    PoseidonTables t;
    if (!poseidon_derive_tables(rc, &t)) { vx_set_error("poseidon: table derivation failed"); return VX_ECUDA; }
    VX_CUDA(poseidon_upload_constants(t, ctx->stream));
    VX_CUDA(cudaMemcpyToSymbolAsync(g_pos_rc, t.rc, sizeof t.rc, 0, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

// ------------------------------------------------------------------------------------------------
GL_D uint64_t pair_pos(uint64_t q, uint32_t lvl) { return 2 * (q * (2ULL << lvl) + (1ULL << lvl) - 1); }

GL_D void store_digest(u64* dst, const u64 s[12]) {
    ulonglong2 a = make_ulonglong2(gl_canon(s[0]), gl_canon(s[1]));
    ulonglong2 b = make_ulonglong2(gl_canon(s[2]), gl_canon(s[3]));
    reinterpret_cast<ulonglong2*>(dst)[0] = a;
    reinterpret_cast<ulonglong2*>(dst)[1] = b;
}

// one thread per leaf: hash_or_noop(leaf) -> interleaved slot (or cap when the subtree is 1 leaf); state in registers.
// LONG: the form for wide leaves on a full GPU -- 256-thread blocks, partial rounds in groups of 4, one barrier per
// permutation to keep the warps of a block in the same code region (see POSEIDON_GROUP).
#ifndef LEAF_LONG_MINB
#define LEAF_LONG_MINB 4
#endif
template <bool COL_MAJOR, bool LONG>
__global__ void __launch_bounds__(LONG ? POSEIDON_BLOCK_LONG : POSEIDON_BLOCK, LONG ? LEAF_LONG_MINB : 1024 / POSEIDON_BLOCK)
leaf_hash_kernel(const u64* __restrict__ leaves, uint64_t stride, uint64_t N, uint32_t c, uint32_t sub_bits,
                 u64* __restrict__ digests, u64* __restrict__ cap) {
    constexpr int BLOCK = LONG ? POSEIDON_BLOCK_LONG : POSEIDON_BLOCK;
    __shared__ u64 scratch[12 * BLOCK];
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= N) return;
    const u64* src = COL_MAJOR ? leaves + row : leaves + row * c;
    const uint64_t step = COL_MAJOR ? stride : 1;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    if (c <= 4) {                         // hash_or_noop: identity, zero padded
#pragma unroll
        for (int i = 0; i < 4; i++) s[i] = ((uint32_t)i < c) ? src[i * step] : 0;
    } else {
        for (uint32_t off = 0; off < c; off += POSEIDON_RATE) {
#pragma unroll
            for (int i = 0; i < POSEIDON_RATE; i++)
                if (off + i < c) s[i] = src[(uint64_t)(off + i) * step];   // overwrite-mode absorb
            if (LONG) {
                __syncthreads();          // exited threads (row >= N) do not take part
                poseidon_permute<POSEIDON_GROUP_LONG, BLOCK>(s, scratch + threadIdx.x);
            } else {
                poseidon_permute(s, scratch + threadIdx.x);
            }
        }
    }
    u64* dst;
    if (sub_bits == 0) {
        dst = cap + 4 * row;
    } else {
        uint64_t sub = 1ULL << sub_bits;
        uint64_t sidx = row >> sub_bits, jj = row & (sub - 1);
        dst = digests + 4 * (sidx * (2 * sub - 2) + 4 * (jj >> 1) + (jj & 1));
    }
    store_digest(dst, s);
}

// Streaming form of the leaf hash for a commit whose columns arrive in chunks (values uploaded from the host while earlier
// columns are already transformed): absorb columns [col0, col1) of every leaf into the sponge state kept in `state`
// (12 x N, lane-major so that a warp's accesses are contiguous).  col0 and every col1 < c are multiples of the rate, so a
// chunk boundary is a permutation boundary of hash_no_pad; the launch with col1 == c writes the digests.
// (the long form carries the sponge state across the launch: 64 registers would spill, so it runs 3 blocks per SM)
template <bool LONG>
__global__ void __launch_bounds__(LONG ? POSEIDON_BLOCK_LONG : POSEIDON_BLOCK, LONG ? 3 : 1024 / POSEIDON_BLOCK)
leaf_absorb_kernel(const u64* __restrict__ lde, uint64_t stride, uint64_t N, uint32_t c, uint32_t col0, uint32_t col1,
                   u64* __restrict__ state, uint32_t sub_bits, u64* __restrict__ digests, u64* __restrict__ cap) {
    constexpr int BLOCK = LONG ? POSEIDON_BLOCK_LONG : POSEIDON_BLOCK;
    __shared__ u64 scratch[12 * BLOCK];
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= N) return;
    const u64* src = lde + row;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = col0 ? state[(uint64_t)i * N + row] : 0;
    for (uint32_t off = col0; off < col1; off += POSEIDON_RATE) {
#pragma unroll
        for (int i = 0; i < POSEIDON_RATE; i++)
            if (off + i < col1) s[i] = src[(uint64_t)(off + i) * stride];
        if (LONG) {
            __syncthreads();
            poseidon_permute<POSEIDON_GROUP_LONG, BLOCK>(s, scratch + threadIdx.x);
        } else {
            poseidon_permute(s, scratch + threadIdx.x);
        }
    }
    if (col1 < c) {
#pragma unroll
        for (int i = 0; i < 12; i++) state[(uint64_t)i * N + row] = s[i];
        return;
    }
    u64* dst;
    if (sub_bits == 0) {
        dst = cap + 4 * row;
    } else {
        uint64_t sub = 1ULL << sub_bits;
        uint64_t sidx = row >> sub_bits, jj = row & (sub - 1);
        dst = digests + 4 * (sidx * (2 * sub - 2) + 4 * (jj >> 1) + (jj & 1));
    }
    store_digest(dst, s);
}

// one thread per sibling pair of layer `lvl`: two_to_one -> parent slot (or cap at the top)
__global__ void __launch_bounds__(POSEIDON_BLOCK) level_hash_kernel(u64* __restrict__ digests, u64* __restrict__ cap,
                                                         uint32_t lvl, uint32_t sub_bits, uint64_t total_pairs) {
    __shared__ u64 scratch[12 * POSEIDON_BLOCK];
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_pairs) return;
    uint32_t pair_bits = sub_bits - lvl - 1;            // pairs per subtree = 2^pair_bits
    uint64_t sidx = t >> pair_bits, q = t & ((1ULL << pair_bits) - 1);
    uint64_t sub = 1ULL << sub_bits;
    u64* blk = digests + 4 * sidx * (2 * sub - 2);
    const ulonglong2* pr = reinterpret_cast<const ulonglong2*>(blk + 4 * pair_pos(q, lvl));
    u64 s[12];
    ulonglong2 v0 = pr[0], v1 = pr[1], v2 = pr[2], v3 = pr[3];
    s[0] = v0.x; s[1] = v0.y; s[2] = v1.x; s[3] = v1.y;
    s[4] = v2.x; s[5] = v2.y; s[6] = v3.x; s[7] = v3.y;
    s[8] = s[9] = s[10] = s[11] = 0;
    poseidon_permute(s, scratch + threadIdx.x);
    u64* dst = (pair_bits == 0) ? cap + 4 * sidx : blk + 4 * (pair_pos(q >> 1, lvl + 1) + (q & 1));
    store_digest(dst, s);
}

// Small levels: the upper levels of the tree have too few pairs to fill the GPU, so their cost is the LATENCY of one
// permutation per level (~44 us with one thread per permutation).  Here 16 lanes cooperate on one two_to_one: lane i
// (< 12) owns state element i, the S-box layer runs in parallel and the MDS layer gathers the 12 elements with warp
// shuffles (spec round structure: add constants, x^7, MDS; partial rounds apply x^7 on lane 0 only).  ~5x lower latency,
// ~3x more thread-instructions: used (inside level_hash_fused_kernel) only while a level has at most VX_COOP_MAX_PAIRS pairs.
// (8192 / 16384 measured: interior levels of 2^19 leaves 0.667 -> 0.691 / 0.751 ms, of a shard's 2^16 leaves 0.390 -> 0.337 /
// 0.361 ms -- profiles/r02_coop_threshold_ab.log; 4096 kept)
#ifndef VX_COOP_MAX_PAIRS
#define VX_COOP_MAX_PAIRS 4096
#endif

// The cooperative permutation: lane (of a 16-lane group) el < 12 holds state element el; returns the permuted element.
// `s` enters WITHOUT the first round's constants.
GL_D u64 poseidon_coop_permute(u64 s, const uint32_t lane, const uint32_t el, const u64* __restrict__ rc) {
    constexpr u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    s = gl_add_canon(s, rc[el]);
    uint32_t from[12];                                               // source lanes of the MDS gather (loop invariant)
#pragma unroll
    for (int i = 0; i < 12; i++) from[i] = el + i >= 12 ? el + i - 12 : el + i;
    const u32 d = el == 0 ? 8u : 0u;
#pragma unroll 1
    for (int r = 0; r < POSEIDON_ROUNDS; r++) {
        const bool full = r < 4 || r >= 26;
        const u64 k = rc[12 * (r + 1) + el];                         // next round's constants (row 30 is zero)
        if (full || lane == 0) s = gl_pow7_cc(s);
        const u32 slo = lo32(s), shi = hi32(s);
        // two independent accumulation chains per half (even / odd terms): the chain of dependent IMAD.WIDE is what
        // bounds the latency of this layer
        u64 al = (u64)lo32(k), ah = (u64)hi32(k);
        u64 bl = mul_wide(slo, d), bh = mul_wide(shi, d);
#pragma unroll
        for (int i = 0; i < 12; i += 2) {
            u32 xl = __shfl_sync(0xffffffffu, slo, from[i], 16);
            u32 xh = __shfl_sync(0xffffffffu, shi, from[i], 16);
            u32 yl = __shfl_sync(0xffffffffu, slo, from[i + 1], 16);
            u32 yh = __shfl_sync(0xffffffffu, shi, from[i + 1], 16);
            al = mad_wide(xl, C[i], al);
            ah = mad_wide(xh, C[i], ah);
            bl = mad_wide(yl, C[i + 1], bl);
            bh = mad_wide(yh, C[i + 1], bh);
        }
        al += bl;                                                    // < 2^42: no carry
        ah += bh;
        u64 l = al + ((u64)lo32(ah) << 32);
        u32 c = l < al;
        s = gl_reduce96_cc(lo32(l), hi32(l), hi32(ah) + c);
    }
    return s;
}

// Several small levels in ONE launch: a CTA owns 2^(nl-1) consecutive sibling pairs of layer lvl0 (all inside one cap
// subtree) and reduces them through nl layers, 16 lanes per two_to_one, handing each layer's digests to the next through
// shared memory (and writing every one to its slot of the interleaved layout).  A standalone launch of a small level
// costs ~24 us, most of it cold instruction / constant caches and launch latency; here only the first layer pays that.
#define VX_FUSE_MAX_LEVELS 7
__global__ void __launch_bounds__(1024) level_hash_fused_kernel(u64* __restrict__ digests, u64* __restrict__ cap,
                                                                uint32_t lvl0, uint32_t nl, uint32_t sub_bits) {
    __shared__ u64 sh[2][4 << (VX_FUSE_MAX_LEVELS - 1)];              // digests of the layer just produced (ping-pong)
    __shared__ u64 sh_rc[31 * 12];                                   // lanes read DIFFERENT constants: a constant-bank read
    for (uint32_t i = threadIdx.x; i < 31 * 12; i += blockDim.x) sh_rc[i] = g_pos_rc[i];     // would serialise 12-fold
    __syncthreads();
    const uint32_t lane = threadIdx.x & 15, g = threadIdx.x >> 4;
    const uint32_t el = lane < 12 ? lane : 0;
    const uint32_t p0 = 1u << (nl - 1);                              // pairs of layer lvl0 owned by this CTA
    const uint32_t pair_bits0 = sub_bits - lvl0 - 1;                 // pairs per subtree at lvl0 = 2^pair_bits0 >= p0
    const uint64_t t0 = (uint64_t)blockIdx.x * p0;                   // first pair (global numbering over subtrees)
    const uint64_t sidx = t0 >> pair_bits0;
    const uint64_t q0 = t0 & ((1ULL << pair_bits0) - 1);
    const uint64_t sub = 1ULL << sub_bits;
    u64* blk = digests + 4 * sidx * (2 * sub - 2);
    for (uint32_t i = 0; i < nl; i++) {
        const uint32_t lvl = lvl0 + i;
        const uint32_t pairs = p0 >> i;
        if ((g & ~1u) < pairs) {                                     // warp-uniform: both groups of a warp shuffle together
            const bool live = g < pairs;
            const uint32_t gg = live ? g : pairs - 1;                // an idle second group redoes the last pair (no store)
            const uint64_t q = (q0 >> i) + gg;
            u64 s = 0;
            if (lane < 8) s = i == 0 ? blk[4 * pair_pos(q, lvl) + lane] : sh[i & 1][8 * gg + lane];
            s = poseidon_coop_permute(s, lane, el, sh_rc);
            if (live && lane < 4) {
                const u64 v = gl_canon(s);
                const bool top = lvl + 1 == sub_bits;
                u64* dst = top ? cap + 4 * sidx : blk + 4 * (pair_pos(q >> 1, lvl + 1) + (q & 1));
                dst[lane] = v;
                sh[(i + 1) & 1][4 * g + lane] = v;
            }
        }
        __syncthreads();
    }
}

// Launch shape of the leaf kernels.  Few leaves (a shard of a sharded commit, the small oracles): 64- or 32-thread blocks
// spread the warps evenly over the SMs (512 blocks of 128 on 148 SMs leave some SMs with 16 warps and others with 12).
// Capping the residency at 28 or 24 warps per SM so that 2^19 leaves become 3.95 whole waves instead of 3.46 was measured
// and does not pay (8.13 / 8.15 / 8.18 ms at 32 / 28 / 24 warps, profiles/r02_poseidon_ab.md).  Neither do, for a shard's 2^16
// leaves (3.5 warps per sub-partition), two half-sponges per leaf block drawn from a ticket counter (1.28 against 1.17 ms) or a
// 128-register build of the kernel (1.17 ms either way): profiles/r02_leaf_split_ab.log.
struct LeafPlan { unsigned threads, blocks; };
static LeafPlan leaf_plan(vx_ctx* ctx, uint64_t N) {
    LeafPlan p;
    p.threads = POSEIDON_BLOCK;
    while (p.threads > 32 && (N + p.threads - 1) / p.threads < 8ULL * (uint64_t)ctx->sm_count) p.threads >>= 1;
    p.blocks = (unsigned)((N + p.threads - 1) / p.threads);
    return p;
}

int32_t merkle_build_device(vx_ctx* ctx, const u64* leaves, bool col_major, uint64_t stride, uint64_t N,
                            uint32_t c, uint32_t cap_height, u64* digests, u64* cap, cudaEvent_t after_leaves) {
    uint32_t log_N = ilog2(N);
    VX_REQUIRE((1ULL << log_N) == N, "merkle: leaf count %llu is not a power of two", (unsigned long long)N);
    VX_REQUIRE(cap_height <= log_N, "merkle: cap_height %u > log2(leaves) %u", cap_height, log_N);
    VX_REQUIRE(c >= 1, "merkle: empty leaves");
    uint32_t sub_bits = log_N - cap_height;
    const LeafPlan lp = leaf_plan(ctx, N);
    // the long form needs a full wave of 256-thread blocks and enough permutations per leaf to amortise its cold start
    const bool long_form = col_major && c >= 64 && lp.threads == POSEIDON_BLOCK &&
                           N / POSEIDON_BLOCK_LONG >= 4ULL * (uint64_t)ctx->sm_count;
    if (long_form)
        leaf_hash_kernel<true, true><<<(unsigned)(N / POSEIDON_BLOCK_LONG), POSEIDON_BLOCK_LONG, 0, ctx->stream>>>(
            leaves, stride, N, c, sub_bits, digests, cap);
    else if (col_major)
        leaf_hash_kernel<true, false><<<lp.blocks, lp.threads, 0, ctx->stream>>>(leaves, stride, N, c, sub_bits, digests, cap);
    else
        leaf_hash_kernel<false, false><<<lp.blocks, lp.threads, 0, ctx->stream>>>(leaves, stride, N, c, sub_bits, digests, cap);
    VX_LAUNCH_COUNT(ctx, 1);
    if (after_leaves) VX_CUDA(cudaEventRecord(after_leaves, ctx->stream));
    return merkle_levels_device(ctx, N, cap_height, digests, cap);
}

int32_t merkle_absorb_device(vx_ctx* ctx, const u64* lde, uint64_t stride, uint64_t N, uint32_t c, uint32_t col0,
                             uint32_t col1, u64* state, uint32_t cap_height, u64* digests, u64* cap) {
    const uint32_t log_N = ilog2(N);
    VX_REQUIRE((1ULL << log_N) == N && cap_height <= log_N, "merkle: bad leaf count / cap height");
    VX_REQUIRE(c > 4 && col0 < col1 && col1 <= c && col0 % POSEIDON_RATE == 0 && (col1 == c || col1 % POSEIDON_RATE == 0),
               "merkle: column chunk [%u, %u) of %u is not aligned to the sponge rate", col0, col1, c);
    const LeafPlan lp = leaf_plan(ctx, N);
    const int slot = ctx->absorb_count < vx_ctx::VX_MAX_ABSORB ? ctx->absorb_count++ : -1;
    if (slot >= 0) VX_CUDA(cudaEventRecord(ctx->absorb_ev[2 * slot], ctx->stream));
    // chunks of 8+ permutations per leaf on a full GPU: the long form of the sponge (see leaf_hash_kernel)
    if (col1 - col0 >= 64 && lp.threads == POSEIDON_BLOCK && N / POSEIDON_BLOCK_LONG >= 3ULL * (uint64_t)ctx->sm_count)
        leaf_absorb_kernel<true><<<(unsigned)(N / POSEIDON_BLOCK_LONG), POSEIDON_BLOCK_LONG, 0, ctx->stream>>>(
            lde, stride, N, c, col0, col1, state, log_N - cap_height, digests, cap);
    else
        leaf_absorb_kernel<false><<<lp.blocks, lp.threads, 0, ctx->stream>>>(
            lde, stride, N, c, col0, col1, state, log_N - cap_height, digests, cap);
    if (slot >= 0) VX_CUDA(cudaEventRecord(ctx->absorb_ev[2 * slot + 1], ctx->stream));
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

// interior levels over leaf digests already in place
int32_t merkle_levels_device(vx_ctx* ctx, uint64_t N, uint32_t cap_height, u64* digests, u64* cap) {
    const uint32_t sub_bits = ilog2(N) - cap_height;
    for (uint32_t lvl = 0; lvl < sub_bits;) {
        uint64_t total_pairs = N >> (lvl + 1);
        if (total_pairs <= VX_COOP_MAX_PAIRS) {
            // wide layers: short runs (many CTAs, spread over the SMs); the narrow top: one long run per subtree group
            uint32_t nl = sub_bits - lvl;
            const uint32_t cap_nl = total_pairs > 256 ? 5 : VX_FUSE_MAX_LEVELS;
            if (nl > cap_nl) nl = cap_nl;
            const uint32_t p0 = 1u << (nl - 1);
            unsigned threads = p0 * 16 < 32 ? 32 : p0 * 16;
            level_hash_fused_kernel<<<(unsigned)(total_pairs / p0), threads, 0, ctx->stream>>>(digests, cap, lvl, nl, sub_bits);
            lvl += nl;
        } else {
            unsigned b = (unsigned)((total_pairs + 127) / 128);
            level_hash_kernel<<<b, 128, 0, ctx->stream>>>(digests, cap, lvl, sub_bits, total_pairs);
            lvl++;
        }
        VX_LAUNCH_COUNT(ctx, 1);
    }
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

// ------------------------------------------------------------------------------------------------ queries
__global__ void merkle_paths_kernel(const u64* __restrict__ digests, uint32_t sub_bits,
                                    const u64* __restrict__ idx, uint32_t k, u64* __restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k * sub_bits) return;
    uint32_t qi = t / sub_bits, lvl = t % sub_bits;
    uint64_t leaf = idx[qi];
    uint64_t sub = 1ULL << sub_bits;
    uint64_t sidx = leaf >> sub_bits, j = leaf & (sub - 1);
    uint64_t node = j >> lvl;
    const u64* src = digests + 4 * (sidx * (2 * sub - 2) + pair_pos(node >> 1, lvl) + ((node & 1) ^ 1));
    u64* dst = out + 4 * ((uint64_t)qi * sub_bits + lvl);
    for (int i = 0; i < 4; i++) dst[i] = src[i];
}

int32_t merkle_paths_device(vx_ctx* ctx, const u64* digests, uint64_t N, uint32_t cap_height,
                            const u64* idx_dev, uint32_t k, u64* siblings_dev) {
    uint32_t sub_bits = ilog2(N) - cap_height;
    if (sub_bits == 0 || k == 0) return VX_OK;
    uint32_t total = k * sub_bits;
    merkle_paths_kernel<<<(total + 127) / 128, 128, 0, ctx->stream>>>(digests, sub_bits, idx_dev, k, siblings_dev);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

__global__ void gather_rows_kernel(const u64* __restrict__ leaves, bool col_major, uint64_t stride, uint32_t c,
                                   const u64* __restrict__ idx, uint32_t k, u64* __restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)k * c) return;
    uint32_t qi = (uint32_t)(t / c), col = (uint32_t)(t % c);
    uint64_t row = idx[qi];
    u64 v = col_major ? leaves[(uint64_t)col * stride + row] : leaves[row * c + col];
    out[t] = gl_canon(v);
}

int32_t gather_rows_device(vx_ctx* ctx, const u64* leaves, bool col_major, uint64_t stride, uint32_t c,
                           const u64* idx_dev, uint32_t k, u64* rows_dev) {
    if (k == 0) return VX_OK;
    uint64_t total = (uint64_t)k * c;
    gather_rows_kernel<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(leaves, col_major, stride, c,
                                                                                 idx_dev, k, rows_dev);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

// column-major (c x stride) -> row-major (N x c), 32x32 tiles through shared memory
__global__ void transpose_rows_kernel(const u64* __restrict__ in, uint64_t stride, uint64_t N, uint32_t c,
                                      u64* __restrict__ out) {
    __shared__ u64 tile[32][33];
    uint64_t row0 = (uint64_t)blockIdx.x * 32;
    uint32_t col0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        uint32_t col = col0 + j;
        uint64_t row = row0 + threadIdx.x;
        if (col < c && row < N) tile[j][threadIdx.x] = in[(uint64_t)col * stride + row];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        uint64_t row = row0 + j;
        uint32_t col = col0 + threadIdx.x;
        if (col < c && row < N) out[row * c + col] = gl_canon(tile[threadIdx.x][j]);
    }
}

int32_t transpose_to_rows_device(vx_ctx* ctx, const u64* colmajor, uint64_t stride, uint64_t N, uint32_t c,
                                 u64* rows_dev) {
    dim3 grid((unsigned)((N + 31) / 32), (c + 31) / 32), block(32, 8);
    transpose_rows_kernel<<<grid, block, 0, ctx->stream>>>(colmajor, stride, N, c, rows_dev);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

// ------------------------------------------------------------------------------------------------ primitives
__global__ void __launch_bounds__(POSEIDON_BLOCK) permute_kernel(const u64* __restrict__ in, uint64_t count, u64* __restrict__ out) {
    __shared__ u64 scratch[12 * POSEIDON_BLOCK];
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = in[t * 12 + i];
    poseidon_permute(s, scratch + threadIdx.x);
#pragma unroll
    for (int i = 0; i < 12; i++) out[t * 12 + i] = gl_canon(s[i]);
}

int32_t poseidon_permute_device(vx_ctx* ctx, const u64* in, uint64_t count, u64* out) {
    if (count == 0) return VX_OK;
    permute_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(in, count, out);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

__global__ void __launch_bounds__(POSEIDON_BLOCK) hash_no_pad_kernel(const u64* __restrict__ in, uint64_t count, uint32_t len,
                                                          u64* __restrict__ out) {
    __shared__ u64 scratch[12 * POSEIDON_BLOCK];
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const u64* src = in + t * len;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    for (uint32_t off = 0; off < len; off += POSEIDON_RATE) {
#pragma unroll
        for (int i = 0; i < POSEIDON_RATE; i++)
            if (off + i < len) s[i] = src[off + i];
        poseidon_permute(s, scratch + threadIdx.x);
    }
    for (int i = 0; i < 4; i++) out[t * 4 + i] = gl_canon(s[i]);
}

int32_t hash_no_pad_device(vx_ctx* ctx, const u64* in, uint64_t count, uint32_t len, u64* out) {
    if (count == 0) return VX_OK;
    hash_no_pad_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(in, count, len, out);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}
