// Shared host-side plumbing for libvectorx_b200: context, error reporting, device buffers.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vectorx_b200.h"
#include "gl.cuh"
#include "twiddle_view.cuh"

// ---- error handling: never throw across the C ABI -----------------------------------------------
void vx_set_error(const char* fmt, ...);

#define VX_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            vx_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? VX_ENOMEM : VX_ECUDA;                  \
        }                                                                                     \
    } while (0)

#define VX_CHECK(expr)                 \
    do {                               \
        int32_t _r = (expr);           \
        if (_r != VX_OK) return _r;    \
    } while (0)

#define VX_REQUIRE(cond, ...)          \
    do {                               \
        if (!(cond)) {                 \
            vx_set_error(__VA_ARGS__); \
            return VX_EINVAL;          \
        }                              \
    } while (0)

// ---- context ----------------------------------------------------------------------------------------
// Twiddle tables (all canonical):
//   w_lo/w_hi : powers of W = POWER_OF_TWO_GENERATOR (order 2^32): W^E = w_hi[E>>16] * w_lo[E&0xffff]
//   wi_lo/wi_hi : same for W^-1
//   g_lo/g_hi : powers of the coset shift g: g^m = g_hi[m>>12] * g_lo[m&0xfff]  (m < 2^24)
//   roots12 / iroots12 : w_4096^e, e < 2048, for the in-shared-memory sub-transforms
// commit phases, bracketed by CUDA events on the context stream
enum { VX_EV_START = 0, VX_EV_STAGED, VX_EV_INTT, VX_EV_LDE, VX_EV_LEAF, VX_EV_TREE, VX_NUM_PHASE_EVENTS };

struct vx_ctx {
    int device = 0;
    int sm_count = 0;
    // commits from host memory travel as column chunks (copy / transform / hash overlap); tuned values, see api.cu
    static constexpr uint32_t h2d_chunks = 8;         // chunks of the non-streamed form (hashers other than Poseidon)
    static constexpr uint32_t stream_chunks = 15;     // upper bound on the chunks of the streamed form
    static constexpr uint32_t h2d_first_groups = 1;   // first streamed chunk in 8-column groups (then doubling)
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; // host->device staging of commit inputs, overlapped with the transforms
    cudaStream_t aux_stream = nullptr;  // producer side of a streamed sharded commit (iNTT + peer push of the own slice)
    // streamed leaf hashing: one event pair per leaf_absorb launch of the most recent commit (their sum is the leaf-hash
    // time vx_ctx_phase_ms reports; the launches interleave with the transforms)
    static constexpr int VX_MAX_ABSORB = 64;
    cudaEvent_t absorb_ev[2 * VX_MAX_ABSORB] = {};
    int absorb_count = 0;
    cudaEvent_t copy_ev[16] = {};       // one per column chunk in flight
    cudaEvent_t copy_free = nullptr;    // staging buffer no longer read by the compute stream
    std::mutex mu;                      // serialises calls on this lane's streams
    std::atomic<uint64_t> launches{0};
    // Lanes: a context is a set of VX_LANES independent stream sets ("lanes") sharing one copy of the tables, so that
    // concurrent callers (Rayon workers, several STARK proofs in flight: SURVEY.md 8b) run side by side instead of queueing
    // on one mutex.  lanes[0] is the context itself; every entry point takes a free lane for the duration of the call.
    vx_ctx* root = nullptr;             // the context the caller holds
    std::vector<vx_ctx*> lanes;         // root only
    std::atomic<unsigned> lane_ticket{0};
    std::atomic<int> last_commit_lane{0};       // lane of the most recent commit (vx_ctx_phase_ms)
    int lane_index = 0;
    u64 *w_lo = nullptr, *w_hi = nullptr, *wi_lo = nullptr, *wi_hi = nullptr;
    u64 *g_lo = nullptr, *g_hi = nullptr, *gi_lo = nullptr, *gi_hi = nullptr;
    u64 *roots12 = nullptr, *iroots12 = nullptr;
    u64 *roots12f = nullptr, *iroots12f = nullptr;      // w_4096^e, e < 4096 (radix-16 group twiddles)
    u64 *inner_fwd = nullptr, *inner_inv = nullptr;     // [16][16] w_256^(+-r bitrev_4(q)): between the two groups of a 256-row pass
    // per-shape derived tables (four-step twiddles of a pass, coset factors of an LDE), built on first use (ntt.cu)
    struct NttTable { uint64_t key; u64* p; size_t bytes; };
    std::vector<NttTable> ntt_cache;
    size_t ntt_cache_bytes = 0;
    std::mutex ntt_cache_mu;
    // phase events of the most recent commit on this context (see vx_ctx_phase_ms)
    cudaEvent_t ev[VX_NUM_PHASE_EVENTS] = {};
};

#define VX_LAUNCH_COUNT(ctx, n) (ctx)->launches.fetch_add((n), std::memory_order_relaxed)

// classify a pointer: returns true if it is device memory
bool vx_is_device_ptr(const void* p);

// RAII device buffer (cudaMallocAsync on the context stream)
struct DevBuf {
    u64* p = nullptr;
    size_t bytes = 0;
    cudaStream_t s = nullptr;
    int32_t alloc(size_t nbytes, cudaStream_t stream);
    void release();
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

static inline unsigned ilog2(uint64_t x) {
    unsigned r = 0;
    while ((1ULL << r) < x) r++;
    return r;
}
#define VX_LANES 4
struct CtxGuard {       // one call at a time on THIS lane (used with the primary lane by the sharded commit, whose group
    std::lock_guard<std::mutex> lk;                                       // state lives on the primary lane's streams)
    explicit CtxGuard(vx_ctx* c) : lk(c->mu) { cudaSetDevice(c->device); }
};
struct LaneGuard {      // takes a free lane of the context (or queues on one, round robin); binds the device to the thread
    vx_ctx* lane = nullptr;
    explicit LaneGuard(vx_ctx* c) {
        vx_ctx* root = c->root ? c->root : c;
        for (vx_ctx* l : root->lanes)
            if (l->mu.try_lock()) { lane = l; break; }
        if (!lane) {
            lane = root->lanes.empty() ? root : root->lanes[root->lane_ticket.fetch_add(1) % root->lanes.size()];
            lane->mu.lock();
        }
        cudaSetDevice(lane->device);
    }
    ~LaneGuard() { lane->mu.unlock(); }
    LaneGuard(const LaneGuard&) = delete;
    LaneGuard& operator=(const LaneGuard&) = delete;
};
// rebinds the local `ctx` to the lane taken for this call
#define VX_LANE(ctx) LaneGuard _lane_guard(ctx); ctx = _lane_guard.lane

// copy helpers that accept host or device memory on either side
static inline int32_t copy_in(vx_ctx* ctx, u64* dst_dev, const u64* src, size_t bytes) {
    VX_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyDefault, ctx->stream));
    return VX_OK;
}
static inline int32_t copy_out(vx_ctx* ctx, u64* dst, const u64* src_dev, size_t bytes) {
    VX_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDefault, ctx->stream));
    return VX_OK;
}

struct vx_batch {
    vx_ctx* ctx;
    uint32_t c, log_n, rate_bits, cap_height;
    uint32_t blk_first, blk_count;      // leaf blocks (cosets) held: all 2^rate_bits unless sharded
    // more shards than cosets: the shard holds sub-block fold_index of the 2^fold_bits equal parts of ONE coset's leaf
    // block (blk_count == 1), computed as a 2^(log_n - fold_bits)-point transform of the folded coefficients (lde_batch)
    uint32_t fold_bits = 0, fold_index = 0;
    uint32_t hasher = 0;                // VX_HASHER_*
    DevBuf coeffs;    // c x n
    DevBuf lde;       // c x N_loc column-major, leaf order
    DevBuf digests;   // 2 (N_loc - caps_loc) x 4
    DevBuf cap;       // caps_loc x 4
    uint64_t n() const { return 1ULL << log_n; }
    uint64_t N() const { return 1ULL << (log_n + rate_bits); }
    uint64_t N_loc() const { return ((uint64_t)blk_count << log_n) >> fold_bits; }
    uint64_t leaf_first() const { return ((uint64_t)blk_first << log_n) + (uint64_t)fold_index * (n() >> fold_bits); }
    uint32_t shard_bits() const { return rate_bits - ilog2(blk_count) + fold_bits; }   // log2(number of shards)
    // shard s of 2^sbits: whole cosets while there are enough of them, then equal parts of one coset
    void set_shard(uint32_t s, uint32_t sbits) {
        if (sbits <= rate_bits) { blk_count = (1u << rate_bits) >> sbits; blk_first = s * blk_count; fold_bits = fold_index = 0; }
        else { fold_bits = sbits - rate_bits; blk_count = 1; blk_first = s >> fold_bits; fold_index = s & ((1u << fold_bits) - 1); }
    }
    uint32_t cap_height_loc() const { return cap_height - shard_bits(); }
};


// ---- module entry points (one per .cu) ------------------------------------------------------------
int32_t poseidon_module_init(vx_ctx* ctx);                 // merkle.cu: uploads round constants
void poseidon_round_constants_host(u64 out[360]);          // merkle.cu: ChaCha8Rng(0) derivation

// merkle.cu ---------------------------------------------------------------------------------------
// Hash N leaves of width c.  col_major: element (row, col) at leaves[col * stride + row];
// otherwise row-major at leaves[row * c + col].  digests: plonky2 interleaved layout (device),
// cap: 2^cap_height x 4 (device).
int32_t merkle_build_device(vx_ctx* ctx, const u64* leaves, bool col_major, uint64_t stride, uint64_t N,
                            uint32_t c, uint32_t cap_height, u64* digests, u64* cap,
                            cudaEvent_t after_leaves = nullptr);
// streaming form: absorb columns [col0, col1) of the column-major leaves into the per-leaf sponge state (12 x N); the
// chunk with col1 == c writes the leaf digests, merkle_levels_device then builds the interior levels
int32_t merkle_absorb_device(vx_ctx* ctx, const u64* lde, uint64_t stride, uint64_t N, uint32_t c, uint32_t col0,
                             uint32_t col1, u64* state, uint32_t cap_height, u64* digests, u64* cap);
int32_t merkle_levels_device(vx_ctx* ctx, uint64_t N, uint32_t cap_height, u64* digests, u64* cap);
int32_t merkle_paths_device(vx_ctx* ctx, const u64* digests, uint64_t N, uint32_t cap_height,
                            const u64* idx_dev, uint32_t k, u64* siblings_dev);
int32_t gather_rows_device(vx_ctx* ctx, const u64* leaves, bool col_major, uint64_t stride, uint32_t c,
                           const u64* idx_dev, uint32_t k, u64* rows_dev);
int32_t transpose_to_rows_device(vx_ctx* ctx, const u64* colmajor, uint64_t stride, uint64_t N, uint32_t c,
                                 u64* rows_dev);
int32_t poseidon_permute_device(vx_ctx* ctx, const u64* in, uint64_t count, u64* out);
int32_t hash_no_pad_device(vx_ctx* ctx, const u64* in, uint64_t count, uint32_t len, u64* out);

// bn128.cu: the wrapper-stage hasher (PoseidonBN128Hash); same digest size and tree layout as above --------------
int32_t bn128_module_init(vx_ctx* ctx);
int32_t bn128_merkle_build_device(vx_ctx* ctx, const u64* leaves, bool col_major, uint64_t stride, uint64_t N,
                                  uint32_t c, uint32_t cap_height, u64* digests, u64* cap,
                                  cudaEvent_t after_leaves = nullptr);
int32_t bn128_permute_device(vx_ctx* ctx, const u64* in, uint64_t count, u64* out);      // count x 4 scalars x 4 words
int32_t bn128_hash_device(vx_ctx* ctx, const u64* in, uint64_t count, uint32_t len, bool or_noop, u64* out);
// dispatch on VX_HASHER_*
static inline int32_t merkle_build_hasher(vx_ctx* ctx, uint32_t hasher, const u64* leaves, bool col_major, uint64_t stride,
                                          uint64_t N, uint32_t c, uint32_t cap_height, u64* digests, u64* cap,
                                          cudaEvent_t after_leaves = nullptr) {
    return hasher == VX_HASHER_POSEIDON_BN128
               ? bn128_merkle_build_device(ctx, leaves, col_major, stride, N, c, cap_height, digests, cap, after_leaves)
               : merkle_build_device(ctx, leaves, col_major, stride, N, c, cap_height, digests, cap, after_leaves);
}

// ntt.cu ------------------------------------------------------------------------------------------
int32_t ntt_module_init(vx_ctx* ctx);
// quotient_jit.cu: the kernel compiled at run time for this circuit's gate program (cudaKernel_t as a launchable handle), or nullptr
const void* quotient_jit_lookup(vx_ctx* ctx, const vx_circuit_desc* d, uint32_t* threads_per_block);
int32_t prover_module_init(vx_ctx* ctx);                   // prover.cu: its own copy of the Poseidon tables
int32_t fri_module_init(vx_ctx* ctx);                      // fri.cu: its own copy of the Poseidon tables
void ntt_module_destroy(vx_ctx* ctx);
// In-place forward DIF NTT of `count` contiguous transforms of size 2^log_n starting at data
// (transform t occupies data[t << log_n ...]); output in bit-reversed order. inverse uses w^-1.
int32_t ntt_dif_inplace(vx_ctx* ctx, u64* data, uint64_t count, uint32_t log_n, bool inverse);
// values (c x n natural) -> coefficients (c x n natural): plonky2 ifft. `work` (c x n) is clobbered.
int32_t intt_batch(vx_ctx* ctx, u64* work, u64* coeffs_out, uint32_t c, uint32_t log_n);
// coefficients (c x n natural) -> LDE on coset g*<w_N>, leaf (bit-reversed) order, c x N column-major.
// Only leaf blocks [blk_first, blk_first + blk_count) of the 2^rate_bits cosets are produced (a block
// is one coset = n leaves); lde_out is c x (blk_count * n).  fold_bits > 0 (then blk_count == 1): only part fold_index
// of the 2^fold_bits equal parts of that block, lde_out is c x (n >> fold_bits).
int32_t lde_batch(vx_ctx* ctx, const u64* coeffs, u64* lde_out, uint32_t c, uint32_t log_n, uint32_t rate_bits,
                  uint32_t blk_first, uint32_t blk_count, uint32_t fold_bits = 0, uint32_t fold_index = 0);
// generic natural-order transform used by vx_ntt (tests, FRI layers)
int32_t ntt_natural(vx_ctx* ctx, const u64* in, u64* out, uint32_t c, uint32_t log_n, bool inverse,
                    uint64_t coset_shift);
// in place: data[col][m] *= shift^m, then DIF NTT -> evaluations on shift*<w_n> in bit-reversed order
int32_t coset_ntt_bitrev_inplace(vx_ctx* ctx, u64* data, uint32_t c, uint32_t log_n, uint64_t shift);
