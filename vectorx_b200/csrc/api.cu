// C ABI of libvectorx_b200 (see include/vectorx_b200.h for what each entry point replaces).
#include "common.cuh"
#include "poseidon_tables.h"

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

void vx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char* vx_last_error(void) { return g_err; }

bool vx_is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int32_t DevBuf::alloc(size_t nbytes, cudaStream_t stream) {
    // Re-allocation of a long-lived buffer (vx_fri's coefficients / values between folds): the old block is freed on the
    // stream of THIS call, i.e. after the kernels just queued there that still read it.  Freeing it on the stream it was
    // allocated on -- another lane's, idle since that call returned -- would hand the block back to the pool while those
    // kernels run (with one stream per context the two were the same stream).
    if (p) cudaFreeAsync(p, stream);
    p = nullptr;
    s = stream;
    bytes = nbytes;
    if (nbytes == 0) return VX_OK;
    cudaError_t e = cudaMallocAsync((void**)&p, nbytes, stream);
    if (e != cudaSuccess) {
        p = nullptr;
        vx_set_error("device allocation of %zu bytes failed: %s", nbytes, cudaGetErrorString(e));
        cudaGetLastError();
        return VX_ENOMEM;
    }
    return VX_OK;
}
void DevBuf::release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
}

// ------------------------------------------------------------------------------------------------ context
// streams and events of one lane
static int32_t lane_create_streams(vx_ctx* l) {
    // the side streams feed the main one (uploads; the producer half of a sharded commit): their kernels go first when
    // the block scheduler has a choice, otherwise a GPU full of hashing CTAs starves the exchange its peers wait for
    int prio_lo = 0, prio_hi = 0;
    VX_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    VX_CUDA(cudaStreamCreateWithPriority(&l->stream, cudaStreamNonBlocking, prio_lo));
    VX_CUDA(cudaStreamCreateWithPriority(&l->copy_stream, cudaStreamNonBlocking, prio_hi));
    VX_CUDA(cudaStreamCreateWithPriority(&l->aux_stream, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < VX_NUM_PHASE_EVENTS; i++) VX_CUDA(cudaEventCreate(&l->ev[i]));
    for (auto& e : l->absorb_ev) VX_CUDA(cudaEventCreate(&e));
    for (auto& e : l->copy_ev) VX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    VX_CUDA(cudaEventCreateWithFlags(&l->copy_free, cudaEventDisableTiming));
    return VX_OK;
}
static void lane_destroy_streams(vx_ctx* l) {
    for (cudaStream_t st : {l->stream, l->copy_stream, l->aux_stream})
        if (st) cudaStreamSynchronize(st);
    for (int i = 0; i < VX_NUM_PHASE_EVENTS; i++) if (l->ev[i]) cudaEventDestroy(l->ev[i]);
    for (auto& e : l->copy_ev) if (e) cudaEventDestroy(e);
    for (auto& e : l->absorb_ev) if (e) cudaEventDestroy(e);
    if (l->copy_free) cudaEventDestroy(l->copy_free);
    for (cudaStream_t st : {l->stream, l->copy_stream, l->aux_stream})
        if (st) cudaStreamDestroy(st);
}

// the one predicate for "this library can run there": the build holds sm_100a code only
static bool device_usable(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return false; }
    return prop.major == 10;
}

extern "C" int32_t vx_ctx_create(int32_t device, vx_ctx** out) {
    if (!out) { vx_set_error("vx_ctx_create: out is NULL"); return VX_EINVAL; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        vx_set_error("vx_ctx_create: no CUDA device visible (there is no CPU fallback)");
        return VX_ENODEV;
    }
    VX_REQUIRE(device >= 0 && device < ndev, "vx_ctx_create: device %d out of range (0..%d)", device, ndev - 1);
    cudaDeviceProp prop;
    VX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (!device_usable(device)) {
        vx_set_error("vx_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                     prop.major, prop.minor);
        return VX_ENODEV;
    }
    VX_CUDA(cudaSetDevice(device));
    vx_ctx* ctx = new (std::nothrow) vx_ctx();
    if (!ctx) return VX_ENOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->root = ctx;
    ctx->lanes.push_back(ctx);
    // keep freed blocks in the stream-ordered pool: commits allocate GBs per call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thresh = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
    int32_t r = lane_create_streams(ctx);
    if (r == VX_OK) r = poseidon_module_init(ctx);
    if (r == VX_OK) r = ntt_module_init(ctx);
    if (r == VX_OK) r = fri_module_init(ctx);
    if (r == VX_OK) r = prover_module_init(ctx);
    if (r == VX_OK) r = bn128_module_init(ctx);
    // the other lanes: own streams and events, the root's tables (read-only after this point)
    for (int i = 1; r == VX_OK && i < VX_LANES; i++) {
        vx_ctx* l = new (std::nothrow) vx_ctx();
        if (!l) { r = VX_ENOMEM; break; }
        l->device = device; l->sm_count = ctx->sm_count; l->root = ctx; l->lane_index = i;
        l->w_lo = ctx->w_lo; l->w_hi = ctx->w_hi; l->wi_lo = ctx->wi_lo; l->wi_hi = ctx->wi_hi;
        l->g_lo = ctx->g_lo; l->g_hi = ctx->g_hi; l->gi_lo = ctx->gi_lo; l->gi_hi = ctx->gi_hi;
        l->roots12 = ctx->roots12; l->iroots12 = ctx->iroots12; l->roots12f = ctx->roots12f; l->iroots12f = ctx->iroots12f;
        l->inner_fwd = ctx->inner_fwd; l->inner_inv = ctx->inner_inv;
        ctx->lanes.push_back(l);
        r = lane_create_streams(l);
    }
    if (r != VX_OK) { vx_ctx_destroy(ctx); return r; }
    *out = ctx;
    return VX_OK;
}

extern "C" void vx_ctx_destroy(vx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (vx_ctx* l : ctx->lanes) lane_destroy_streams(l);
    ntt_module_destroy(ctx);
    for (size_t i = 1; i < ctx->lanes.size(); i++) delete ctx->lanes[i];
    delete ctx;
}

extern "C" int32_t vx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int usable = 0;
    for (int d = 0; d < n; d++) usable += device_usable(d) ? 1 : 0;
    return usable;
}
// CUDA ordinals of the devices vx_ctx_create accepts (a box may mix GPU generations); returns how many were written
extern "C" int32_t vx_device_list(int32_t* ordinals_out, int32_t capacity) {
    int n = 0, k = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (int d = 0; d < n; d++)
        if (device_usable(d)) {
            if (ordinals_out && k < capacity) ordinals_out[k] = d;
            k++;
        }
    return k;
}
extern "C" int32_t vx_device_sync(vx_ctx* ctx) {
    VX_REQUIRE(ctx, "vx_device_sync: ctx is NULL");
    for (vx_ctx* l : ctx->root->lanes)
        for (cudaStream_t st : {l->stream, l->copy_stream, l->aux_stream}) VX_CUDA(cudaStreamSynchronize(st));
    return VX_OK;
}
extern "C" int32_t vx_ctx_phase_ms(vx_ctx* ctx, float out[VX_NUM_PHASE_EVENTS - 1]) {
    VX_REQUIRE(ctx && out, "vx_ctx_phase_ms: NULL argument");
    ctx = ctx->root->lanes[ctx->root->last_commit_lane.load()];          // the lane of the most recent commit
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (int i = 0; i + 1 < VX_NUM_PHASE_EVENTS; i++) {
        out[i] = 0.f;
        cudaError_t e = cudaEventElapsedTime(&out[i], ctx->ev[i], ctx->ev[i + 1]);
        if (e != cudaSuccess) { cudaGetLastError(); out[i] = -1.f; }
    }
    // streamed commits hash inside the "lde" bracket: move the time of the leaf_absorb launches to "leaf_hash"
    float leaf = 0.f;
    for (int k = 0; k < ctx->absorb_count; k++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ctx->absorb_ev[2 * k], ctx->absorb_ev[2 * k + 1]) == cudaSuccess) leaf += t;
        else cudaGetLastError();
    }
    if (ctx->absorb_count && out[VX_EV_INTT] >= 0.f && out[VX_EV_LDE] >= 0.f) {
        out[VX_EV_INTT] -= leaf;            // bracket INTT -> LDE ("lde")
        out[VX_EV_LDE] += leaf;             // bracket LDE -> LEAF ("leaf_hash")
    }
    return VX_OK;
}
extern "C" void* vx_ctx_stream(vx_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t vx_ctx_launch_count(vx_ctx* ctx) {
    if (!ctx) return 0;
    uint64_t n = 0;
    for (vx_ctx* l : ctx->root->lanes) n += l->launches.load();
    return n;
}

extern "C" int32_t vx_dev_alloc(vx_ctx* ctx, size_t bytes, uint64_t** out) {
    VX_REQUIRE(ctx && out, "vx_dev_alloc: NULL argument");
    *out = nullptr;
    VX_LANE(ctx);
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 8, ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        vx_set_error("vx_dev_alloc: %zu bytes: %s", bytes, cudaGetErrorString(e));
        return VX_ENOMEM;
    }
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = (uint64_t*)p;
    return VX_OK;
}
extern "C" void vx_dev_free(vx_ctx* ctx, uint64_t* p) {
    if (!ctx || !p) return;
    VX_LANE(ctx);
    cudaFreeAsync(p, ctx->stream);
}
extern "C" int32_t vx_dev_copy(vx_ctx* ctx, void* dst, const void* src, size_t bytes) {
    VX_REQUIRE(ctx && (bytes == 0 || (dst && src)), "vx_dev_copy: NULL argument");
    if (bytes == 0) return VX_OK;
    VX_LANE(ctx);
    VX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

extern "C" int32_t vx_host_alloc(size_t bytes, void** out) {
    VX_REQUIRE(out, "vx_host_alloc: NULL argument");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 8, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        vx_set_error("vx_host_alloc: %zu bytes: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? VX_ENOMEM : VX_ECUDA;
    }
    return VX_OK;
}
extern "C" void vx_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess) cudaGetLastError();
}
extern "C" int32_t vx_host_register(void* p, size_t bytes) {
    VX_REQUIRE(p && bytes, "vx_host_register: NULL argument");
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        vx_set_error("vx_host_register: %zu bytes at %p: %s", bytes, p, cudaGetErrorString(e));
        return VX_ECUDA;
    }
    return VX_OK;
}
extern "C" void vx_host_unregister(void* p) {
    if (p && cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
}

#define EV(ctx, i) VX_CUDA(cudaEventRecord((ctx)->ev[i], (ctx)->stream))

// Where a commit's input lives: ONE contiguous c x n matrix (`flat`), or c separately allocated columns (`cols`, what
// plonky2 holds: Vec<PolynomialValues<F>>, one heap allocation per column).  Host memory may be pageable, pinned or
// registered; device memory is accepted too.
struct ColSource {
    const u64* flat = nullptr;
    const u64* const* cols = nullptr;
    const u64* col(uint32_t j, uint64_t n) const { return cols ? cols[j] : flat + (size_t)j * n; }
};
// 0 = device, 1 = pinned / registered host, 2 = pageable host
static int mem_kind(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 2; }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return 0;
    return a.type == cudaMemoryTypeHost ? 1 : 2;
}
// copies columns [c0, c1) to dst (contiguous): one transfer for a flat source, one per column otherwise
static int32_t copy_columns(const ColSource& src, uint32_t c0, uint32_t c1, uint64_t n, u64* dst, cudaMemcpyKind kind,
                            cudaStream_t st) {
    if (src.flat) {
        VX_CUDA(cudaMemcpyAsync(dst, src.flat + (size_t)c0 * n, (size_t)(c1 - c0) * n * sizeof(u64), kind, st));
        return VX_OK;
    }
    for (uint32_t j = c0; j < c1; j++)
        VX_CUDA(cudaMemcpyAsync(dst + (size_t)(j - c0) * n, src.cols[j], n * sizeof(u64), kind, st));
    return VX_OK;
}

static int32_t commit_run(vx_ctx* ctx, vx_batch* b, const ColSource& src, bool is_values, u64* keep = nullptr) {
    const uint32_t c = b->c;
    const uint64_t n = b->n(), N_loc = b->N_loc();
    const size_t coeff_bytes = (size_t)c * n * sizeof(u64);
    const uint64_t caps_loc = 1ULL << b->cap_height_loc();
    VX_CHECK(b->coeffs.alloc(coeff_bytes, ctx->stream));
    VX_CHECK(b->lde.alloc((size_t)c * N_loc * sizeof(u64), ctx->stream));
    VX_CHECK(b->digests.alloc((size_t)2 * (N_loc - caps_loc) * 4 * sizeof(u64), ctx->stream));
    VX_CHECK(b->cap.alloc((size_t)caps_loc * 4 * sizeof(u64), ctx->stream));
    EV(ctx, VX_EV_START);
    ctx->absorb_count = 0;
    const int kind = mem_kind(src.col(0, n));
    const bool from_host = kind != 0;
    // column chunks: with a host source the H2D copy of chunk k+1 (copy stream) overlaps the transforms of chunk k
    const uint32_t nchunks = (from_host && c >= 16) ? ctx->h2d_chunks : 1;
    DevBuf stage;
    u64* work = nullptr;
    if (is_values) {
        // stage the values in the (not yet used) LDE buffer when it is big enough and chunks cannot collide with it
        work = b->lde.p;
        if (nchunks > 1 || b->lde.bytes < coeff_bytes) { VX_CHECK(stage.alloc(coeff_bytes, ctx->stream)); work = stage.p; }
    }
    if (nchunks > 1) {
        VX_CUDA(cudaEventRecord(ctx->copy_free, ctx->stream));             // allocations above are stream-ordered
        VX_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_free, 0));
    }
    // streaming sponge: chunk boundaries fall on multiples of the sponge rate and every chunk is absorbed into the per-leaf
    // state right after its LDE, so only the LAST chunk's hashing is left when the last copy lands
    const bool stream = nchunks > 1 && b->hasher == VX_HASHER_POSEIDON && c > 4;
    // Chunks grow geometrically (8, 16, 32, ... columns) so that the copy of chunk k+1 is never longer than the work on
    // the chunks before it; the doubling stops once the rest can hide too.  Work per column / copy per column is about
    // 2^rate_bits (hashing is per LDE row, the copy per trace row): ~8 at rate_bits 3 -- 2^16 x 135 goes as 8 / 16 / 111
    // columns -- and ~2 at rate_bits 1, where a 2502-column STARK trace needs nine chunks.
    const uint32_t groups = (c + 7) / 8;
    uint32_t bound[17] = {0};
    uint32_t nchunks_eff = nchunks;
    // Pageable memory travels through the driver's staging buffers at roughly a seventh of the pinned rate (~8 GB/s measured
    // against ~55): the ratio drops accordingly, chunks grow by it (not by 2) and, where the copy cannot hide at all
    // (ratio < 1), stay equal in size.  Measured at 2^16 x 135, rate_bits 3: 10.96 ms per commit from 135 pageable columns
    // (9.87 resident, 10.21 pinned).  A library-side staging ring filled by four host threads was measured too: 10.80 ms,
    // but it competes with the caller's own threads for cores (batch_prove on two ranks: 76 -> 50 proofs/s) and is gone;
    // callers that want PCIe rate register their Vecs (vx_host_register).
    if (stream) {
#ifndef VX_PAGEABLE_SLOWDOWN
#define VX_PAGEABLE_SLOWDOWN 7.0
#endif
        const double ratio = ((double)(1u << b->rate_bits) + 0.3) * (kind == 2 ? 1.0 / VX_PAGEABLE_SLOWDOWN : 1.0);
        const double grow = ratio >= 2.0 ? 2.0 : (ratio > 1.0 ? ratio : 1.0);
        uint32_t k = 0, g = 0;
        double size = ctx->h2d_first_groups;
        if (grow == 1.0 && size * (ctx->stream_chunks - 1) < groups) size = (double)((groups + ctx->stream_chunks - 2) / (ctx->stream_chunks - 1));
        while (k + 1 < ctx->stream_chunks && g + (uint32_t)size < groups) {
            g += (uint32_t)size;
            bound[++k] = 8 * g;
            size *= grow;
            if ((double)(groups - g) <= 0.6 * ratio * (double)g) break;      // the rest hides behind what is queued
        }
        bound[++k] = c;
        nchunks_eff = k;
    }
    DevBuf sponge;
    if (stream) VX_CHECK(sponge.alloc((size_t)12 * N_loc * sizeof(u64), ctx->stream));
    for (uint32_t k = 0; k < nchunks_eff; k++) {
        uint32_t c0 = (uint32_t)((uint64_t)c * k / nchunks), c1 = (uint32_t)((uint64_t)c * (k + 1) / nchunks);
        if (stream) { c0 = bound[k]; c1 = bound[k + 1]; }
        if (c1 == c0) continue;
        const size_t off = (size_t)c0 * n, bytes = (size_t)(c1 - c0) * n * sizeof(u64);
        u64* dst = is_values ? work + off : b->coeffs.p + off;
        // `keep` (vx_commit_from_values_keep): the chunk lands in the caller's device copy of the values first and is
        // duplicated into the work buffer on the device (the inverse transform runs in place)
        u64* land = keep ? keep + off : dst;
        if (nchunks > 1) {
            VX_CHECK(copy_columns(src, c0, c1, n, land, cudaMemcpyHostToDevice, ctx->copy_stream));
            VX_CUDA(cudaEventRecord(ctx->copy_ev[k], ctx->copy_stream));
            VX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[k], 0));
        } else {
            VX_CHECK(copy_columns(src, c0, c1, n, land, cudaMemcpyDefault, ctx->stream));
        }
        if (keep) VX_CUDA(cudaMemcpyAsync(dst, land, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        if (nchunks == 1) EV(ctx, VX_EV_STAGED);
        if (is_values) VX_CHECK(intt_batch(ctx, work + off, b->coeffs.p + off, c1 - c0, b->log_n));
        if (nchunks == 1) EV(ctx, VX_EV_INTT);
        VX_CHECK(lde_batch(ctx, b->coeffs.p + off, b->lde.p + (size_t)c0 * N_loc, c1 - c0, b->log_n, b->rate_bits,
                           b->blk_first, b->blk_count, b->fold_bits, b->fold_index));
        if (stream)
            VX_CHECK(merkle_absorb_device(ctx, b->lde.p, N_loc, N_loc, c, c0, c1, sponge.p, b->cap_height_loc(),
                                          b->digests.p, b->cap.p));
    }
    if (nchunks > 1) { EV(ctx, VX_EV_STAGED); EV(ctx, VX_EV_INTT); }      // phases interleave: all reported under "lde"
    EV(ctx, VX_EV_LDE);
    if (stream) {
        EV(ctx, VX_EV_LEAF);
        VX_CHECK(merkle_levels_device(ctx, N_loc, b->cap_height_loc(), b->digests.p, b->cap.p));
    } else {
        VX_CHECK(merkle_build_hasher(ctx, b->hasher, b->lde.p, true, N_loc, N_loc, c, b->cap_height_loc(), b->digests.p,
                                     b->cap.p, ctx->ev[VX_EV_LEAF]));
    }
    EV(ctx, VX_EV_TREE);
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

static ColSource flat_src(const uint64_t* p) { ColSource s; s.flat = (const u64*)p; return s; }

static int32_t commit_impl(vx_ctx* ctx, const ColSource& src, bool is_values, uint32_t c, uint32_t log_n,
                           uint32_t rate_bits, uint32_t cap_height, uint32_t shard_index, uint32_t shard_count,
                           vx_batch** out, uint32_t hasher = VX_HASHER_POSEIDON, u64* keep = nullptr) {
    VX_REQUIRE(ctx && (src.flat || src.cols) && out, "commit: NULL argument");
    *out = nullptr;
    if (src.cols)
        for (uint32_t j = 0; j < c; j++) VX_REQUIRE(src.cols[j], "commit: column pointer %u is NULL", j);
    VX_REQUIRE(hasher <= VX_HASHER_POSEIDON_BN128, "commit: unknown hasher %u", hasher);
    VX_REQUIRE(c >= 1 && c < 16384, "commit: column count %u out of range", c);
    VX_REQUIRE(log_n <= 26 && log_n + rate_bits <= 30, "commit: 2^%u rows at rate_bits %u unsupported (coset tables cover 2^26 rows)", log_n, rate_bits);
    VX_REQUIRE(cap_height <= log_n + rate_bits, "commit: cap_height %u exceeds tree height %u", cap_height,
               log_n + rate_bits);
    VX_REQUIRE(shard_count >= 1 && (shard_count & (shard_count - 1)) == 0 && shard_index < shard_count,
               "commit: bad shard %u of %u", shard_index, shard_count);
    uint32_t sbits = ilog2(shard_count);
    VX_REQUIRE(sbits <= cap_height && sbits <= rate_bits + log_n,
               "commit: %u shards need cap_height >= %u (whole cap subtrees per shard)", shard_count, sbits);
    VX_LANE(ctx);
    vx_batch* b = new (std::nothrow) vx_batch();
    if (!b) return VX_ENOMEM;
    b->ctx = ctx->root; b->c = c; b->log_n = log_n; b->rate_bits = rate_bits; b->cap_height = cap_height;
    b->hasher = hasher;
    b->set_shard(shard_index, sbits);
    ctx->root->last_commit_lane.store(ctx->lane_index);
    int32_t r = commit_run(ctx, b, src, is_values, keep);
    if (r != VX_OK) {
        // a chunked upload may still be in flight on the copy stream, targeting buffers the batch is about to free
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->aux_stream);
        cudaStreamSynchronize(ctx->stream);
        delete b;
        return r;
    }
    *out = b;
    return VX_OK;
}

extern "C" int32_t vx_commit_from_values(vx_ctx* ctx, const uint64_t* cols, uint32_t c, uint32_t log_n,
                                         uint32_t rate_bits, uint32_t cap_height, vx_batch** out) {
    return commit_impl(ctx, flat_src(cols), true, c, log_n, rate_bits, cap_height, 0, 1, out);
}
extern "C" int32_t vx_commit_from_values_keep(vx_ctx* ctx, const uint64_t* cols, uint32_t c, uint32_t log_n,
                                              uint32_t rate_bits, uint32_t cap_height, uint64_t* values_dev_out,
                                              vx_batch** out) {
    VX_REQUIRE(values_dev_out && vx_is_device_ptr(values_dev_out), "vx_commit_from_values_keep: values_dev_out must be device memory");
    return commit_impl(ctx, flat_src(cols), true, c, log_n, rate_bits, cap_height, 0, 1, out, VX_HASHER_POSEIDON,
                       (u64*)values_dev_out);
}
// plonky2's own input shape: c separately allocated columns (Vec<PolynomialValues<F>>), no flattening on the host
extern "C" int32_t vx_commit_from_values_cols(vx_ctx* ctx, const uint64_t* const* cols, uint32_t c, uint32_t log_n,
                                              uint32_t rate_bits, uint32_t cap_height, vx_batch** out) {
    ColSource s;
    s.cols = (const u64* const*)cols;
    return commit_impl(ctx, s, true, c, log_n, rate_bits, cap_height, 0, 1, out);
}
extern "C" int32_t vx_commit_from_coeffs_cols(vx_ctx* ctx, const uint64_t* const* coeffs, uint32_t c, uint32_t log_n,
                                              uint32_t rate_bits, uint32_t cap_height, vx_batch** out) {
    ColSource s;
    s.cols = (const u64* const*)coeffs;
    return commit_impl(ctx, s, false, c, log_n, rate_bits, cap_height, 0, 1, out);
}
extern "C" int32_t vx_commit_from_coeffs(vx_ctx* ctx, const uint64_t* coeffs, uint32_t c, uint32_t log_n,
                                         uint32_t rate_bits, uint32_t cap_height, vx_batch** out) {
    return commit_impl(ctx, flat_src(coeffs), false, c, log_n, rate_bits, cap_height, 0, 1, out);
}
extern "C" int32_t vx_commit_from_values_hasher(vx_ctx* ctx, uint32_t hasher, const uint64_t* cols, uint32_t c,
                                                uint32_t log_n, uint32_t rate_bits, uint32_t cap_height, vx_batch** out) {
    return commit_impl(ctx, flat_src(cols), true, c, log_n, rate_bits, cap_height, 0, 1, out, hasher);
}
extern "C" int32_t vx_commit_from_coeffs_hasher(vx_ctx* ctx, uint32_t hasher, const uint64_t* coeffs, uint32_t c,
                                                uint32_t log_n, uint32_t rate_bits, uint32_t cap_height, vx_batch** out) {
    return commit_impl(ctx, flat_src(coeffs), false, c, log_n, rate_bits, cap_height, 0, 1, out, hasher);
}
extern "C" int32_t vx_commit_from_coeffs_shard(vx_ctx* ctx, const uint64_t* coeffs, uint32_t c, uint32_t log_n,
                                               uint32_t rate_bits, uint32_t cap_height, uint32_t shard_index,
                                               uint32_t shard_count, vx_batch** out) {
    return commit_impl(ctx, flat_src(coeffs), false, c, log_n, rate_bits, cap_height, shard_index, shard_count, out);
}

extern "C" void vx_batch_free(vx_batch* b) {
    if (!b) return;
    {
        LaneGuard g(b->ctx);
        b->coeffs.release(); b->lde.release(); b->digests.release(); b->cap.release();
    }
    delete b;
}

extern "C" int32_t vx_batch_shape(const vx_batch* b, uint32_t out[4]) {
    VX_REQUIRE(b && out, "vx_batch_shape: NULL argument");
    out[0] = b->c; out[1] = b->log_n; out[2] = b->rate_bits; out[3] = b->cap_height;
    return VX_OK;
}

extern "C" int32_t vx_batch_shard(const vx_batch* b, uint64_t out[3]) {
    VX_REQUIRE(b && out, "vx_batch_shard: NULL argument");
    out[0] = b->leaf_first(); out[1] = b->N_loc(); out[2] = 1ULL << b->cap_height_loc();
    return VX_OK;
}

extern "C" int32_t vx_batch_cap(vx_batch* b, uint64_t* cap_out) {
    VX_REQUIRE(b && cap_out, "vx_batch_cap: NULL argument");
    vx_ctx* ctx = b->ctx;
    VX_LANE(ctx);
    VX_CHECK(copy_out(ctx, (u64*)cap_out, b->cap.p, b->cap.bytes));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

extern "C" int32_t vx_batch_coeffs(vx_batch* b, uint64_t* coeffs_out) {
    VX_REQUIRE(b && coeffs_out, "vx_batch_coeffs: NULL argument");
    vx_ctx* ctx = b->ctx;
    VX_LANE(ctx);
    VX_CHECK(copy_out(ctx, (u64*)coeffs_out, b->coeffs.p, b->coeffs.bytes));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

static int32_t check_indices(const uint64_t* idx, uint32_t k, uint64_t N, const char* who) {
    for (uint32_t i = 0; i < k; i++)
        VX_REQUIRE(idx[i] < N, "%s: leaf index %llu out of range (%llu leaves)", who,
                   (unsigned long long)idx[i], (unsigned long long)N);
    return VX_OK;
}

static int32_t query_rows(vx_ctx* ctx, const u64* leaves, bool col_major, uint64_t stride, uint32_t c, uint64_t N,
                          const uint64_t* idx, uint32_t k, uint64_t* rows_out) {
    if (k == 0) return VX_OK;
    VX_CHECK(check_indices(idx, k, N, "leaves"));
    DevBuf di, dr;
    VX_CHECK(di.alloc(k * sizeof(u64), ctx->stream));
    VX_CHECK(dr.alloc((size_t)k * c * sizeof(u64), ctx->stream));
    VX_CHECK(copy_in(ctx, di.p, (const u64*)idx, k * sizeof(u64)));
    VX_CHECK(gather_rows_device(ctx, leaves, col_major, stride, c, di.p, k, dr.p));
    VX_CHECK(copy_out(ctx, (u64*)rows_out, dr.p, dr.bytes));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

static int32_t query_paths(vx_ctx* ctx, const u64* digests, uint64_t N, uint32_t cap_height, const uint64_t* idx,
                           uint32_t k, uint64_t* siblings_out) {
    uint32_t depth = ilog2(N) - cap_height;
    if (k == 0 || depth == 0) return VX_OK;
    VX_CHECK(check_indices(idx, k, N, "merkle_paths"));
    DevBuf di, ds;
    VX_CHECK(di.alloc(k * sizeof(u64), ctx->stream));
    VX_CHECK(ds.alloc((size_t)k * depth * 4 * sizeof(u64), ctx->stream));
    VX_CHECK(copy_in(ctx, di.p, (const u64*)idx, k * sizeof(u64)));
    VX_CHECK(merkle_paths_device(ctx, digests, N, cap_height, di.p, k, ds.p));
    VX_CHECK(copy_out(ctx, (u64*)siblings_out, ds.p, ds.bytes));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

extern "C" int32_t vx_batch_leaves(vx_batch* b, const uint64_t* idx, uint32_t k, uint64_t* rows_out) {
    VX_REQUIRE(b && (k == 0 || (idx && rows_out)), "vx_batch_leaves: NULL argument");
    vx_ctx* ctx = b->ctx;
    VX_LANE(ctx);
    return query_rows(ctx, b->lde.p, true, b->N_loc(), b->c, b->N_loc(), idx, k, rows_out);
}

extern "C" int32_t vx_batch_merkle_paths(vx_batch* b, const uint64_t* idx, uint32_t k, uint64_t* siblings_out) {
    VX_REQUIRE(b && (k == 0 || (idx && siblings_out)), "vx_batch_merkle_paths: NULL argument");
    vx_ctx* ctx = b->ctx;
    VX_LANE(ctx);
    return query_paths(ctx, b->digests.p, b->N_loc(), b->cap_height_loc(), idx, k, siblings_out);
}

extern "C" int32_t vx_batch_download(vx_batch* b, uint64_t* leaves_out, uint64_t* digests_out) {
    VX_REQUIRE(b, "vx_batch_download: NULL batch");
    vx_ctx* ctx = b->ctx;
    VX_LANE(ctx);
    if (leaves_out) {
        DevBuf rows;
        VX_CHECK(rows.alloc(b->lde.bytes, ctx->stream));
        VX_CHECK(transpose_to_rows_device(ctx, b->lde.p, b->N_loc(), b->N_loc(), b->c, rows.p));
        VX_CHECK(copy_out(ctx, (u64*)leaves_out, rows.p, rows.bytes));
        VX_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (digests_out && b->digests.bytes) {
        VX_CHECK(copy_out(ctx, (u64*)digests_out, b->digests.p, b->digests.bytes));
        VX_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return VX_OK;
}

extern "C" const uint64_t* vx_batch_lde_device(const vx_batch* b) { return b ? (const uint64_t*)b->lde.p : nullptr; }
extern "C" const uint64_t* vx_batch_coeffs_device(const vx_batch* b) { return b ? (const uint64_t*)b->coeffs.p : nullptr; }
extern "C" const uint64_t* vx_batch_digests_device(const vx_batch* b) { return b ? (const uint64_t*)b->digests.p : nullptr; }

// ------------------------------------------------------------------------------------------------ MerkleTree
struct vx_tree {
    vx_ctx* ctx;
    uint64_t n;
    uint32_t w, cap_height;
    DevBuf leaves, digests, cap;
};

extern "C" int32_t vx_merkle_new_hasher(vx_ctx* ctx, uint32_t hasher, const uint64_t* leaves, uint64_t n, uint32_t w,
                                        uint32_t cap_height, uint64_t* digests_out, uint64_t* cap_out, vx_tree** tree_out) {
    VX_REQUIRE(ctx && leaves, "vx_merkle_new: NULL argument");
    if (tree_out) *tree_out = nullptr;
    VX_REQUIRE(hasher <= VX_HASHER_POSEIDON_BN128, "vx_merkle_new: unknown hasher %u", hasher);
    VX_REQUIRE(n >= 1 && (n & (n - 1)) == 0, "vx_merkle_new: leaf count %llu is not a power of two",
               (unsigned long long)n);
    VX_REQUIRE(w >= 1, "vx_merkle_new: empty leaves");
    VX_REQUIRE(cap_height <= ilog2(n), "vx_merkle_new: cap_height %u exceeds tree height %u", cap_height, ilog2(n));
    VX_LANE(ctx);
    vx_tree* t = new (std::nothrow) vx_tree();
    if (!t) return VX_ENOMEM;
    t->ctx = ctx->root; t->n = n; t->w = w; t->cap_height = cap_height;
    int32_t r = t->leaves.alloc((size_t)n * w * sizeof(u64), ctx->stream);
    if (r == VX_OK) r = t->digests.alloc((size_t)2 * (n - (1ULL << cap_height)) * 4 * sizeof(u64), ctx->stream);
    if (r == VX_OK) r = t->cap.alloc((size_t)(1ULL << cap_height) * 4 * sizeof(u64), ctx->stream);
    if (r == VX_OK) r = copy_in(ctx, t->leaves.p, (const u64*)leaves, t->leaves.bytes);
    if (r == VX_OK) r = merkle_build_hasher(ctx, hasher, t->leaves.p, false, 0, n, w, cap_height, t->digests.p, t->cap.p);
    if (r == VX_OK && digests_out && t->digests.bytes) r = copy_out(ctx, (u64*)digests_out, t->digests.p, t->digests.bytes);
    if (r == VX_OK && cap_out) r = copy_out(ctx, (u64*)cap_out, t->cap.p, t->cap.bytes);
    if (r == VX_OK) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { vx_set_error("vx_merkle_new: %s", cudaGetErrorString(e)); r = VX_ECUDA; }
    }
    if (r != VX_OK || !tree_out) {
        t->leaves.release(); t->digests.release(); t->cap.release();
        delete t;
        return r;
    }
    *tree_out = t;
    return VX_OK;
}

extern "C" int32_t vx_merkle_new(vx_ctx* ctx, const uint64_t* leaves, uint64_t n, uint32_t w, uint32_t cap_height,
                                 uint64_t* digests_out, uint64_t* cap_out, vx_tree** tree_out) {
    return vx_merkle_new_hasher(ctx, VX_HASHER_POSEIDON, leaves, n, w, cap_height, digests_out, cap_out, tree_out);
}

extern "C" int32_t vx_tree_prove(vx_tree* t, const uint64_t* idx, uint32_t k, uint64_t* siblings_out) {
    VX_REQUIRE(t && (k == 0 || (idx && siblings_out)), "vx_tree_prove: NULL argument");
    vx_ctx* ctx = t->ctx;
    VX_LANE(ctx);
    return query_paths(ctx, t->digests.p, t->n, t->cap_height, idx, k, siblings_out);
}
extern "C" int32_t vx_tree_leaves(vx_tree* t, const uint64_t* idx, uint32_t k, uint64_t* rows_out) {
    VX_REQUIRE(t && (k == 0 || (idx && rows_out)), "vx_tree_leaves: NULL argument");
    vx_ctx* ctx = t->ctx;
    VX_LANE(ctx);
    return query_rows(ctx, t->leaves.p, false, 0, t->w, t->n, idx, k, rows_out);
}
extern "C" int32_t vx_tree_cap(vx_tree* t, uint64_t* cap_out) {
    VX_REQUIRE(t && cap_out, "vx_tree_cap: NULL argument");
    vx_ctx* ctx = t->ctx;
    VX_LANE(ctx);
    VX_CHECK(copy_out(ctx, (u64*)cap_out, t->cap.p, t->cap.bytes));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}
extern "C" void vx_tree_free(vx_tree* t) {
    if (!t) return;
    {
        LaneGuard g(t->ctx);
        t->leaves.release(); t->digests.release(); t->cap.release();
    }
    delete t;
}

// ------------------------------------------------------------------------------------------------ primitives
extern "C" int32_t vx_poseidon_permute(vx_ctx* ctx, const uint64_t* in, uint64_t count, uint64_t* out) {
    VX_REQUIRE(ctx && (count == 0 || (in && out)), "vx_poseidon_permute: NULL argument");
    if (count == 0) return VX_OK;
    VX_LANE(ctx);
    DevBuf a, b;
    VX_CHECK(a.alloc(count * 12 * sizeof(u64), ctx->stream));
    VX_CHECK(b.alloc(count * 12 * sizeof(u64), ctx->stream));
    VX_CHECK(copy_in(ctx, a.p, (const u64*)in, a.bytes));
    VX_CHECK(poseidon_permute_device(ctx, a.p, count, b.p));
    VX_CHECK(copy_out(ctx, (u64*)out, b.p, b.bytes));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

extern "C" int32_t vx_hash_no_pad(vx_ctx* ctx, const uint64_t* in, uint64_t count, uint32_t len, uint64_t* out) {
    VX_REQUIRE(ctx && (count == 0 || out), "vx_hash_no_pad: NULL argument");
    VX_REQUIRE(len == 0 || in, "vx_hash_no_pad: NULL input");
    if (count == 0) return VX_OK;
    VX_LANE(ctx);
    DevBuf a, b;
    VX_CHECK(a.alloc((size_t)count * len * sizeof(u64), ctx->stream));
    VX_CHECK(b.alloc(count * 4 * sizeof(u64), ctx->stream));
    if (len) VX_CHECK(copy_in(ctx, a.p, (const u64*)in, a.bytes));
    VX_CHECK(hash_no_pad_device(ctx, a.p, count, len, b.p));
    VX_CHECK(copy_out(ctx, (u64*)out, b.p, b.bytes));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

// ---- PoseidonBN128Hash primitives (bn128.cu)
static bool bn128_words_canonical(const uint64_t* w) {
    static const uint64_t N[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    for (int i = 3; i >= 0; i--)
        if (w[i] != N[i]) return w[i] < N[i];
    return false;
}
extern "C" int32_t vx_bn128_permute(vx_ctx* ctx, const uint64_t* in, uint64_t count, uint64_t* out) {
    VX_REQUIRE(ctx && (count == 0 || (in && out)), "vx_bn128_permute: NULL argument");
    if (count == 0) return VX_OK;
    for (uint64_t i = 0; i < 4 * count; i++)
        VX_REQUIRE(bn128_words_canonical(in + 4 * i), "vx_bn128_permute: scalar %llu is not below the BN254 scalar modulus",
                   (unsigned long long)i);
    VX_LANE(ctx);
    DevBuf a, b;
    VX_CHECK(a.alloc(count * 16 * sizeof(u64), ctx->stream));
    VX_CHECK(b.alloc(count * 16 * sizeof(u64), ctx->stream));
    VX_CHECK(copy_in(ctx, a.p, (const u64*)in, a.bytes));
    VX_CHECK(bn128_permute_device(ctx, a.p, count, b.p));
    VX_CHECK(copy_out(ctx, (u64*)out, b.p, b.bytes));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}
extern "C" int32_t vx_bn128_hash(vx_ctx* ctx, const uint64_t* in, uint64_t count, uint32_t len, int32_t or_noop,
                                 uint64_t* out) {
    VX_REQUIRE(ctx && (count == 0 || out), "vx_bn128_hash: NULL argument");
    VX_REQUIRE(len == 0 || in, "vx_bn128_hash: NULL input");
    if (count == 0) return VX_OK;
    VX_LANE(ctx);
    DevBuf a, b;
    VX_CHECK(a.alloc((size_t)count * len * sizeof(u64), ctx->stream));
    VX_CHECK(b.alloc(count * 4 * sizeof(u64), ctx->stream));
    if (len) VX_CHECK(copy_in(ctx, a.p, (const u64*)in, a.bytes));
    VX_CHECK(bn128_hash_device(ctx, a.p, count, len, or_noop != 0, b.p));
    VX_CHECK(copy_out(ctx, (u64*)out, b.p, b.bytes));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

// host-side scalar permutation for the Fiat-Shamir challenger (spec form: add constants, x^7, MDS)
extern "C" int32_t vx_challenger_permute(uint64_t state[12]) {
    if (!state) { vx_set_error("vx_challenger_permute: NULL"); return VX_EINVAL; }
    static u64 rc[360];
    static std::once_flag once;
    std::call_once(once, [] { poseidon_round_constants_host(rc); });
    static const u64 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u64 s[12];
    for (int i = 0; i < 12; i++) s[i] = state[i] % GL_P;
    for (int r = 0; r < 30; r++) {
        const bool full = r < 4 || r >= 26;
        for (int i = 0; i < 12; i++) {
            u64 x = s[i] + rc[12 * r + i];
            if (x < s[i] || x >= GL_P) x -= GL_P;
            if (full || i == 0) {
                u64 x2 = gl_mul_slow(x, x), x4 = gl_mul_slow(x2, x2), x3 = gl_mul_slow(x2, x);
                x = gl_mul_slow(x3, x4);
            }
            s[i] = x;
        }
        u64 t[12];
        for (int j = 0; j < 12; j++) {
            unsigned __int128 acc = (j == 0) ? (unsigned __int128)8 * s[0] : 0;
            for (int i = 0; i < 12; i++) acc += (unsigned __int128)s[(i + j) % 12] * CIRC[i];
            t[j] = (u64)(acc % GL_P);
        }
        memcpy(s, t, sizeof s);
    }
    memcpy(state, s, sizeof s);
    return VX_OK;
}

extern "C" int32_t vx_poseidon_constants(uint64_t out[360]) {
    if (!out) { vx_set_error("vx_poseidon_constants: NULL"); return VX_EINVAL; }
    poseidon_round_constants_host((u64*)out);
    return VX_OK;
}

extern "C" int32_t vx_poseidon_fast_tables(uint64_t dense_d[144], uint64_t dense_e[12], uint64_t k[22], uint64_t v[242],
                                           uint64_t w[242]) {
    if (!dense_d || !dense_e || !k || !v || !w) { vx_set_error("vx_poseidon_fast_tables: NULL"); return VX_EINVAL; }
    u64 rc[360];
    poseidon_round_constants_host(rc);
    PoseidonTables t;
    if (!poseidon_derive_tables(rc, &t)) { vx_set_error("poseidon: table derivation failed"); return VX_EINVAL; }
    for (int i = 0; i < 144; i++) dense_d[i] = t.dense_d[i];
    for (int i = 0; i < 12; i++) dense_e[i] = t.dense_e[i];
    for (int i = 0; i < 22; i++) k[i] = t.pk[i];
    for (int i = 0; i < 242; i++) { v[i] = t.pv[i]; w[i] = t.pw[i]; }
    return VX_OK;
}

extern "C" int32_t vx_ntt(vx_ctx* ctx, const uint64_t* in, uint64_t* out, uint32_t c, uint32_t log_n,
                          int32_t inverse, uint64_t coset_shift) {
    VX_REQUIRE(ctx && in && out, "vx_ntt: NULL argument");
    VX_REQUIRE(c >= 1 && log_n <= 26, "vx_ntt: shape out of range");
    VX_LANE(ctx);
    size_t bytes = ((size_t)c << log_n) * sizeof(u64);
    DevBuf a, b;
    VX_CHECK(a.alloc(bytes, ctx->stream));
    VX_CHECK(b.alloc(bytes, ctx->stream));
    VX_CHECK(copy_in(ctx, a.p, (const u64*)in, bytes));
    VX_CHECK(ntt_natural(ctx, a.p, b.p, c, log_n, inverse != 0, coset_shift));
    VX_CHECK(copy_out(ctx, (u64*)out, b.p, bytes));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}
