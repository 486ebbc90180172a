// One commit sharded over G GPUs of an NVSwitch box (coset partition), with the exchange step done by
// this library's own kernels over peer memory -- no NCCL on the data path.
//
// Replaces (and extends to several devices) plonky2 v0.2.0 `PolynomialBatch::from_values`, reached in the reference
// from contracts/lib/succinctx/plonky2x/core/src/backend/circuit/build.rs:69-75; upstream has no multi-device form.
//
//   rank r:  iNTT of its column slice  ->  peer_push_kernel: the coefficient block is stored into EVERY rank's
//            gather buffer with 128-bit stores over NVLink and the last CTA raises this rank's flag on every peer
//            (release, system scope)  ->  coset NTTs of its own columns while the other blocks are in flight, then
//            peer_wait_kernel spins (acquire) until the other ranks' flags carry this commit's epoch and their columns
//            are extended  ->  leaf hashing + cap subtrees of the rank's own leaf block (no communication)
//            ->  the 2^cap/G local cap entries are pushed to every peer the same way.
//
// Every rank ends with the full cap and its own shard (vx_batch) of the commitment.  The gather buffers, the cap
// buffer and the flags of a rank live in ONE cudaMalloc allocation that peers map either directly (same process,
// cudaDeviceEnablePeerAccess) or through a CUDA IPC handle (one process per GPU).
//
// Reuse of the buffers by consecutive commits is ordered by the protocol itself: a rank pushes the coefficients of
// commit k+1 only after it has seen every rank's cap flag of commit k, and a rank raises that flag only after it has
// finished reading its gather buffer.  A wait that does not complete within ~2 s (a peer died) sets an error word
// instead of hanging the device.
#include "common.cuh"

#define VX_MAX_SHARDS 16
#define VX_FLAG_PHASES 16          // 0: whole coefficient slice, 1: cap entries, 2 + j: sub-block j of a streamed commit
#define VX_MAX_SUB 12              // sub-blocks per rank slice in the streamed form (flag phases 2 .. 2 + VX_MAX_SUB - 1)

struct PeerTable {
    u64* base[VX_MAX_SHARDS];
};

struct vx_shard_group {
    vx_ctx* ctx = nullptr;
    uint32_t rank = 0, world = 1, c = 0, log_n = 0, rate_bits = 0, cap_height = 0;
    uint32_t cpr = 0;                  // columns per rank in the gather buffer (last rank zero-padded)
    size_t cap_off = 0, flag_off = 0;  // in u64 from base
    size_t bytes = 0;
    u64* base = nullptr;               // this rank's allocation
    PeerTable peers{};                 // every rank's allocation as mapped into this process
    bool ipc_opened[VX_MAX_SHARDS] = {};
    bool connected = false;
    uint32_t epoch = 0;
    cudaEvent_t push_ev[VX_MAX_SUB] = {};      // streamed form: sub-block j of the OWN slice has been pushed (aux stream)
    long long timeout_cycles = 4000000000LL;   // a peer wait gives up after this many SM cycles (~2 s)
    bool poisoned = false;             // a wait timed out: flags / epochs of the group are no longer consistent
    int layout_mode = 0;               // 0 = by size, 1 = contiguous column slices, 2 = interleaved (see group_interleaved)
    uint32_t* counter = nullptr;       // device: CTAs done (last-CTA election of the push kernel)
    int* err_host = nullptr;           // pinned + mapped: set by a wait that timed out
    int* err_dev = nullptr;
};

// ------------------------------------------------------------------------------------------------ kernels
GL_D void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
GL_D uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// src: count16 x 16 bytes (contiguous).  Every rank g receives the block at peers.base[g] + dst_off.  When the whole
// grid has stored, flag word `flag_index` (u32 units from flag base) on every rank is set to `epoch`.
__global__ void __launch_bounds__(256) peer_push_kernel(const ulonglong2* __restrict__ src, uint64_t count16,
                                                         const __grid_constant__ PeerTable peers, uint32_t world,
                                                         uint64_t dst_off, uint64_t flag_off, uint32_t flag_index,
                                                         uint32_t epoch, uint32_t* __restrict__ counter) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count16; i += stride) {
        const ulonglong2 v = src[i];
#pragma unroll 1
        for (uint32_t g = 0; g < world; g++) reinterpret_cast<ulonglong2*>(peers.base[g] + dst_off)[i] = v;
    }
    __threadfence_system();                       // this thread's peer stores are performed before the election below
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    if (threadIdx.x == 0) *counter = 0;
    if (threadIdx.x < world)
        st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x] + flag_off) + flag_index, epoch);
}

// <<<1, world>>>: thread g waits until rank g's flag reaches `epoch`
__global__ void peer_wait_kernel(const uint32_t* __restrict__ flags, uint32_t epoch, long long timeout_cycles,
                                 int* __restrict__ err) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
        if (clock64() - t0 > timeout_cycles) {
            *err = 1 + (int)threadIdx.x;
            __threadfence_system();
            break;
        }
        __nanosleep(100);
    }
}

// ------------------------------------------------------------------------------------------------ group
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" int32_t vx_shard_group_create(vx_ctx* ctx, uint32_t rank, uint32_t world, uint32_t c, uint32_t log_n,
                                         uint32_t rate_bits, uint32_t cap_height, vx_shard_group** out) {
    VX_REQUIRE(ctx && out, "vx_shard_group_create: NULL argument");
    *out = nullptr;
    VX_REQUIRE(world >= 1 && world <= VX_MAX_SHARDS && (world & (world - 1)) == 0 && rank < world,
               "vx_shard_group_create: bad rank %u of %u", rank, world);
    const uint32_t sbits = ilog2(world);
    VX_REQUIRE(sbits <= cap_height && sbits <= rate_bits + log_n,
               "vx_shard_group_create: %u shards need cap_height >= %u (whole cap subtrees per shard)", world, sbits);
    VX_REQUIRE(c >= 1 && c < 16384 && log_n >= 1 && log_n <= 26 && log_n + rate_bits <= 30 && cap_height <= log_n + rate_bits,
               "vx_shard_group_create: shape out of range");
    CtxGuard g(ctx);
    vx_shard_group* s = new (std::nothrow) vx_shard_group();
    if (!s) return VX_ENOMEM;
    s->ctx = ctx; s->rank = rank; s->world = world; s->c = c; s->log_n = log_n; s->rate_bits = rate_bits;
    s->cap_height = cap_height;
    s->cpr = (c + world - 1) / world;
    const size_t n = (size_t)1 << log_n;
    s->cap_off = align_up((size_t)world * s->cpr * n, 32);
    s->flag_off = align_up(s->cap_off + ((size_t)4 << cap_height), 32);
    s->bytes = (s->flag_off + VX_FLAG_PHASES * VX_MAX_SHARDS / 2) * sizeof(u64);    // u32 flags: [phase][rank]
    cudaError_t e = cudaMalloc((void**)&s->base, s->bytes);          // plain cudaMalloc: exportable through CUDA IPC
    if (e != cudaSuccess) {
        cudaGetLastError();
        vx_set_error("vx_shard_group_create: %zu bytes: %s", s->bytes, cudaGetErrorString(e));
        delete s;
        return VX_ENOMEM;
    }
    bool ok = cudaMemsetAsync(s->base, 0, s->bytes, ctx->stream) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&s->counter, sizeof(uint32_t)) == cudaSuccess;
    ok = ok && cudaMemsetAsync(s->counter, 0, sizeof(uint32_t), ctx->stream) == cudaSuccess;
    ok = ok && cudaHostAlloc((void**)&s->err_host, sizeof(int), cudaHostAllocMapped) == cudaSuccess;
    if (ok) {
        *s->err_host = 0;
        ok = cudaHostGetDevicePointer((void**)&s->err_dev, s->err_host, 0) == cudaSuccess;
    }
    for (auto& e : s->push_ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    if (!ok) {
        vx_set_error("vx_shard_group_create: %s", cudaGetErrorString(cudaGetLastError()));
        for (auto& e : s->push_ev) if (e) cudaEventDestroy(e);
        if (s->counter) cudaFree(s->counter);
        if (s->err_host) cudaFreeHost(s->err_host);
        cudaFree(s->base);
        delete s;
        return VX_ECUDA;
    }
    s->peers.base[rank] = s->base;
    s->connected = world == 1;
    *out = s;
    return VX_OK;
}

extern "C" int32_t vx_shard_group_ipc_handle(vx_shard_group* s, uint8_t handle_out[64]) {
    VX_REQUIRE(s && handle_out, "vx_shard_group_ipc_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    CtxGuard g(s->ctx);
    cudaIpcMemHandle_t h;
    VX_CUDA(cudaIpcGetMemHandle(&h, s->base));
    memcpy(handle_out, &h, 64);
    return VX_OK;
}

// handles: world x 64 bytes, entry r = rank r's vx_shard_group_ipc_handle (one process per GPU)
extern "C" int32_t vx_shard_group_connect_ipc(vx_shard_group* s, const uint8_t* handles) {
    VX_REQUIRE(s && handles, "vx_shard_group_connect_ipc: NULL argument");
    CtxGuard g(s->ctx);
    for (uint32_t r = 0; r < s->world; r++) {
        if (r == s->rank || s->ipc_opened[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * r, 64);
        void* p = nullptr;
        VX_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->peers.base[r] = (u64*)p;
        s->ipc_opened[r] = true;
    }
    s->connected = true;
    return VX_OK;
}

// all ranks live in this process (one context per device): map each other's allocations directly
extern "C" int32_t vx_shard_group_connect_local(vx_shard_group* const* groups, uint32_t world) {
    VX_REQUIRE(groups && world >= 1 && world <= VX_MAX_SHARDS, "vx_shard_group_connect_local: bad argument");
    for (uint32_t r = 0; r < world; r++)
        VX_REQUIRE(groups[r] && groups[r]->world == world && groups[r]->rank == r,
                   "vx_shard_group_connect_local: groups[%u] is not rank %u of %u", r, r, world);
    for (uint32_t a = 0; a < world; a++) {
        vx_shard_group* s = groups[a];
        CtxGuard g(s->ctx);
        for (uint32_t b = 0; b < world; b++) {
            if (a == b) continue;
            const int peer_dev = groups[b]->ctx->device;
            if (peer_dev != s->ctx->device) {
                int can = 0;
                VX_CUDA(cudaDeviceCanAccessPeer(&can, s->ctx->device, peer_dev));
                VX_REQUIRE(can, "vx_shard_group_connect_local: device %d cannot access device %d", s->ctx->device, peer_dev);
                cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) VX_CUDA(e);
                cudaGetLastError();
            }
            s->peers.base[b] = groups[b]->base;
        }
        s->connected = true;
    }
    return VX_OK;
}

extern "C" void vx_shard_group_free(vx_shard_group* s) {
    if (!s) return;
    {
        CtxGuard g(s->ctx);
        cudaStreamSynchronize(s->ctx->stream);
        for (uint32_t r = 0; r < s->world; r++)
            if (s->ipc_opened[r]) cudaIpcCloseMemHandle(s->peers.base[r]);
        if (s->counter) cudaFree(s->counter);
        if (s->err_host) cudaFreeHost(s->err_host);
        if (s->base) cudaFree(s->base);
        for (auto& e : s->push_ev) if (e) cudaEventDestroy(e);
    }
    delete s;
}

/* how long a rank waits for a peer's contribution before the commit fails (default ~2 s); ms == 0 restores the default */
extern "C" int32_t vx_shard_group_set_timeout(vx_shard_group* s, uint32_t ms) {
    VX_REQUIRE(s, "vx_shard_group_set_timeout: NULL argument");
    s->timeout_cycles = ms ? (long long)ms * 2000000LL : 4000000000LL;       // ~2 GHz SM clock
    return VX_OK;
}

/* this rank's gather buffer (world * cols_per_rank x n coefficients, valid after a commit) */
extern "C" const uint64_t* vx_shard_group_coeffs_device(const vx_shard_group* s) { return s ? (const uint64_t*)s->base : nullptr; }
extern "C" uint32_t vx_shard_group_cols_per_rank(const vx_shard_group* s) { return s ? s->cpr : 0; }

// ------------------------------------------------------------------------------------------------ commit
static int32_t push_block(vx_shard_group* s, cudaStream_t stream, const u64* src, uint64_t count_u64, uint64_t dst_off,
                          uint32_t phase) {
    vx_ctx* ctx = s->ctx;
    const uint64_t count16 = count_u64 / 2;
    uint64_t blocks = (count16 + 255) / 256;
    const uint64_t max_blocks = (uint64_t)ctx->sm_count * 4;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks == 0) blocks = 1;
    peer_push_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const ulonglong2*>(src), count16, s->peers,
                                                                 s->world, dst_off, s->flag_off,
                                                                 phase * VX_MAX_SHARDS + s->rank, s->epoch, s->counter);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}
// wait for the flags of ranks [first, first + count) of `phase`
static int32_t wait_flags(vx_shard_group* s, uint32_t phase, uint32_t first, uint32_t count) {
    vx_ctx* ctx = s->ctx;
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(s->base + s->flag_off) + phase * VX_MAX_SHARDS + first;
    peer_wait_kernel<<<1, count, 0, ctx->stream>>>(flags, s->epoch, s->timeout_cycles, s->err_dev);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

#define EV(ctx, i) VX_CUDA(cudaEventRecord((ctx)->ev[i], (ctx)->stream))
#define VX_STREAM_EXCHANGE_BYTES (256ULL << 20)

// ---- which columns a rank transforms and pushes (the iNTT stage) ------------------------------------------------------
// Contiguous (small slices): rank q owns global columns [q cpr, (q + 1) cpr).  The leaf sponge absorbs columns in global
// order, so every rank needs rank 0's slice first and rank 0's NVLink egress paces everybody: fine when the exchange is
// microseconds, not when it is milliseconds (a 2^18 x 2502 trace on 8 ranks: 4.6 GB per rank).
// Interleaved (big slices): every slice is cut into the same <= VX_MAX_SUB parts and part j of rank q is global columns
// [G e_j + q w_j, G e_j + (q + 1) w_j) (e_j = first local column of part j, w_j its width).  In global order the parts of
// all ranks alternate, every link carries what every consumer needs next, and the exchange streams behind the hashing.
static bool group_big(const vx_shard_group* s) {
    return (uint64_t)s->cpr * ((uint64_t)1 << s->log_n) * sizeof(u64) * (s->world - 1) >= VX_STREAM_EXCHANGE_BYTES;
}
static bool group_interleaved(const vx_shard_group* s) {
    return s->world > 1 && (s->layout_mode == 2 || (s->layout_mode == 0 && group_big(s)));
}
// local column edges of the parts of rank q's slice; returns the number of parts
static uint32_t group_sub_edges(const vx_shard_group* s, uint32_t q, uint32_t* sub) {
    uint32_t k = 0;
    sub[0] = 0;
    if (group_interleaved(s)) {
        const uint32_t part = ((s->cpr + VX_MAX_SUB - 1) / VX_MAX_SUB + 7) / 8 * 8;
        for (uint32_t edge = part; k + 1 < VX_MAX_SUB && edge < s->cpr; edge += part) sub[++k] = edge;
    } else if (q == 0) {
        // only the slice that comes first in sponge order is needed early: 8 / 16 / the rest columns
        for (uint32_t edge = 8, step = 16; k + 1 < 3 && edge < s->cpr; edge += step, step *= 2) sub[++k] = edge;
    }
    sub[++k] = s->cpr;
    return k;
}
// first global column of part j of rank q
static uint32_t group_global_col(const vx_shard_group* s, uint32_t q, const uint32_t* sub, uint32_t j) {
    return group_interleaved(s) ? s->world * sub[j] + q * (sub[j + 1] - sub[j]) : q * s->cpr + sub[j];
}

/* 0 = choose by size (default), 1 = contiguous column slices, 2 = interleaved parts; every rank of a group must use the
 * same value, set before the first commit */
extern "C" int32_t vx_shard_group_set_layout(vx_shard_group* s, uint32_t mode) {
    VX_REQUIRE(s && mode <= 2, "vx_shard_group_set_layout: bad argument");
    s->layout_mode = (int)mode;
    return VX_OK;
}
/* global column index of each of this rank's cols_per_rank local columns (UINT32_MAX: padding beyond c): which columns of
 * the trace the caller places in `values_local` */
extern "C" int32_t vx_shard_group_column_map(const vx_shard_group* s, uint32_t* global_col_out) {
    VX_REQUIRE(s && global_col_out, "vx_shard_group_column_map: NULL argument");
    uint32_t sub[VX_MAX_SUB + 1];
    const uint32_t nsub = group_sub_edges(s, s->rank, sub);
    for (uint32_t j = 0; j < nsub; j++) {
        const uint32_t g0 = group_global_col(s, s->rank, sub, j);
        for (uint32_t i = sub[j]; i < sub[j + 1]; i++) {
            const uint32_t g = g0 + (i - sub[j]);
            global_col_out[i] = g < s->c ? g : UINT32_MAX;
        }
    }
    return VX_OK;
}

// Streamed form (the default): the rank's slice is split into sub-blocks of 8 / 16 / the rest columns.
//   producer (copy stream + aux stream): copy sub-block j in -> iNTT -> push to every rank's gather buffer, flag (2 + j, rank)
//   consumer (context stream): global columns in sponge order -- for rank q = 0..G-1, sub-block j: wait for its flag,
//            extend its columns (own cosets), and absorb every complete group of 8 columns into the leaf sponge.
// The first rank's first 8 columns reach every consumer after one small copy + transform + push; from then on copies,
// exchange and transforms hide behind the hashing.  No rank's producer ever waits for a peer, so every flag is raised.
static int32_t shard_commit_run_stream(vx_shard_group* s, vx_batch* b, const u64* values_local, u64* cap_all_out) {
    vx_ctx* ctx = s->ctx;
    const uint64_t n = b->n(), N_loc = b->N_loc();
    const uint64_t caps_loc = 1ULL << b->cap_height_loc();
    const size_t slice_bytes = (size_t)s->cpr * n * sizeof(u64);
    VX_CHECK(b->coeffs.alloc((size_t)s->c * n * sizeof(u64), ctx->stream));
    VX_CHECK(b->lde.alloc((size_t)s->c * N_loc * sizeof(u64), ctx->stream));
    VX_CHECK(b->digests.alloc((size_t)2 * (N_loc - caps_loc) * 4 * sizeof(u64), ctx->stream));
    VX_CHECK(b->cap.alloc((size_t)caps_loc * 4 * sizeof(u64), ctx->stream));
    DevBuf work, mine, sponge;
    VX_CHECK(work.alloc(slice_bytes, ctx->stream));
    VX_CHECK(mine.alloc(slice_bytes, ctx->stream));
    VX_CHECK(sponge.alloc((size_t)12 * N_loc * sizeof(u64), ctx->stream));
    EV(ctx, VX_EV_START);
    ctx->absorb_count = 0;
    uint32_t sub[VX_MAX_SUB + 1];
    const uint32_t nsub = group_sub_edges(s, s->rank, sub);
    const bool inter = group_interleaved(s);
    // ---- producer
    VX_CUDA(cudaEventRecord(ctx->copy_free, ctx->stream));                 // allocations above are stream-ordered
    VX_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_free, 0));
    VX_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->copy_free, 0));
    cudaStream_t main_stream = ctx->stream;
    for (uint32_t j = 0; j < nsub; j++) {
        const size_t off = (size_t)sub[j] * n, cnt = (size_t)(sub[j + 1] - sub[j]) * n;
        VX_CUDA(cudaMemcpyAsync(work.p + off, values_local + off, cnt * sizeof(u64), cudaMemcpyDefault, ctx->copy_stream));
        VX_CUDA(cudaEventRecord(ctx->copy_ev[j], ctx->copy_stream));
        VX_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->copy_ev[j], 0));
        ctx->stream = ctx->aux_stream;                                      // the transforms enqueue on ctx->stream
        int32_t r = intt_batch(ctx, work.p + off, mine.p + off, sub[j + 1] - sub[j], b->log_n);
        ctx->stream = main_stream;
        VX_CHECK(r);
        VX_CHECK(push_block(s, ctx->aux_stream, mine.p + off, cnt, (uint64_t)s->rank * s->cpr * n + off, 2 + j));
        VX_CUDA(cudaEventRecord(s->push_ev[j], ctx->aux_stream));
    }
    VX_CUDA(cudaEventRecord(ctx->copy_ev[VX_MAX_SUB], ctx->aux_stream));  // `work` / `mine` free after this
    EV(ctx, VX_EV_STAGED);
    EV(ctx, VX_EV_INTT);
    // ---- consumer: parts in GLOBAL column order (contiguous: rank-major; interleaved: part-major)
    uint32_t absorbed = 0;
    uint32_t qsub[VX_MAX_SHARDS][VX_MAX_SUB + 1], nq[VX_MAX_SHARDS], max_parts = 0;
    for (uint32_t q = 0; q < s->world; q++) {
        nq[q] = group_sub_edges(s, q, qsub[q]);
        if (nq[q] > max_parts) max_parts = nq[q];
    }
    const uint32_t outer = inter ? max_parts : s->world, inner = inter ? s->world : max_parts;
    for (uint32_t a = 0; a < outer; a++) {
        for (uint32_t bb = 0; bb < inner; bb++) {
            const uint32_t q = inter ? bb : a, j = inter ? a : bb;
            if (j >= nq[q]) continue;
            const uint32_t g0 = group_global_col(s, q, qsub[q], j), g1 = g0 + (qsub[q][j + 1] - qsub[q][j]);
            const uint32_t c0 = g0 < s->c ? g0 : s->c, c1 = g1 < s->c ? g1 : s->c;
            // the own slice is ordered by an event (a spin on a flag raised by another stream of the SAME device would need
            // the two streams to run concurrently, which serialising tools -- ncu, compute-sanitizer -- do not grant)
            if (q == s->rank) VX_CUDA(cudaStreamWaitEvent(ctx->stream, s->push_ev[j], 0));
            else VX_CHECK(wait_flags(s, 2 + j, q, 1));                      // also orders the reuse of the gather buffer
            if (c0 >= c1) continue;
            const u64* src = s->base + ((size_t)q * s->cpr + qsub[q][j]) * n;
            VX_CHECK(lde_batch(ctx, src, b->lde.p + (size_t)c0 * N_loc, c1 - c0, b->log_n, b->rate_bits,
                               b->blk_first, b->blk_count, b->fold_bits, b->fold_index));
            VX_CUDA(cudaMemcpyAsync(b->coeffs.p + (size_t)c0 * n, src, (size_t)(c1 - c0) * n * sizeof(u64),
                                    cudaMemcpyDeviceToDevice, ctx->stream));
            const uint32_t upto = c1 == s->c ? s->c : (c1 / 8) * 8;
            if (upto > absorbed) {
                VX_CHECK(merkle_absorb_device(ctx, b->lde.p, N_loc, N_loc, s->c, absorbed, upto, sponge.p, b->cap_height_loc(),
                                              b->digests.p, b->cap.p));
                absorbed = upto;
            }
        }
    }
    EV(ctx, VX_EV_LDE);
    EV(ctx, VX_EV_LEAF);
    VX_CHECK(merkle_levels_device(ctx, N_loc, b->cap_height_loc(), b->digests.p, b->cap.p));
    EV(ctx, VX_EV_TREE);
    VX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[VX_MAX_SUB], 0));
    VX_CHECK(push_block(s, ctx->stream, b->cap.p, caps_loc * 4, s->cap_off + (uint64_t)s->rank * caps_loc * 4, 1));
    VX_CHECK(wait_flags(s, 1, 0, s->world));
    if (cap_all_out)
        VX_CHECK(copy_out(ctx, cap_all_out, s->base + s->cap_off, ((size_t)4 << s->cap_height) * sizeof(u64)));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->aux_stream));
    if (*s->err_host) {
        vx_set_error("vx_shard_commit_from_values: timed out waiting for rank %d (peer missing or failed); the group's flags "
                     "are no longer consistent: free it and create a new one", *s->err_host - 1);
        *s->err_host = 0;
        s->poisoned = true;
        return VX_ECUDA;
    }
    return VX_OK;
}

static int32_t shard_commit_run(vx_shard_group* s, vx_batch* b, const u64* values_local, u64* cap_all_out) {
    vx_ctx* ctx = s->ctx;
    // When does the column pipeline pay?  Values in host memory on 2 ranks: the copy of a 68-column slice hides behind the
    // hashing, 6.58 -> 6.25 ms end to end.  On 4 / 8 ranks the slices are short, a rank hashes only N/4 / N/8 leaves per
    // column and the extra launches cost more than the copy they hide (3.64 -> 3.71 ms, 2.55 -> 2.79 ms); values already
    // in HBM: whole-slice launches are faster (5.92 against 6.05 ms on 2 ranks).
    // Big slices (>= VX_STREAM_EXCHANGE_BYTES received per rank, e.g. 4.6 GB for a 2^18 x 2502 trace on 8 ranks): the
    // exchange takes milliseconds and only the pipeline hides it, whatever the rank count and wherever the values live.
    const bool want_stream = (s->world <= 2 && !vx_is_device_ptr(values_local)) || group_interleaved(s);
    if (want_stream && s->c > 4) return shard_commit_run_stream(s, b, values_local, cap_all_out);
    ctx->absorb_count = 0;
    const uint64_t n = b->n(), N_loc = b->N_loc();
    const uint64_t caps_loc = 1ULL << b->cap_height_loc();
    const size_t slice_bytes = (size_t)s->cpr * n * sizeof(u64);
    VX_CHECK(b->coeffs.alloc((size_t)s->c * n * sizeof(u64), ctx->stream));
    VX_CHECK(b->lde.alloc((size_t)s->c * N_loc * sizeof(u64), ctx->stream));
    VX_CHECK(b->digests.alloc((size_t)2 * (N_loc - caps_loc) * 4 * sizeof(u64), ctx->stream));
    VX_CHECK(b->cap.alloc((size_t)caps_loc * 4 * sizeof(u64), ctx->stream));
    DevBuf work, mine;
    VX_CHECK(work.alloc(slice_bytes, ctx->stream));
    VX_CHECK(mine.alloc(slice_bytes, ctx->stream));
    EV(ctx, VX_EV_START);
    VX_CHECK(copy_in(ctx, work.p, values_local, slice_bytes));
    EV(ctx, VX_EV_STAGED);
    VX_CHECK(intt_batch(ctx, work.p, mine.p, s->cpr, b->log_n));
    // exchange: my coefficient block -> every rank's gather buffer, on the side stream so that it overlaps the LDE of
    // my own columns (read straight from `mine`)
    VX_CUDA(cudaEventRecord(ctx->copy_free, ctx->stream));
    VX_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_free, 0));
    VX_CHECK(push_block(s, ctx->copy_stream, mine.p, (uint64_t)s->cpr * n, (uint64_t)s->rank * s->cpr * n, 0));
    VX_CUDA(cudaEventRecord(ctx->copy_ev[0], ctx->copy_stream));
    EV(ctx, VX_EV_INTT);
    // the LDE is per column: own columns first, then the columns of the ranks above and below (two contiguous ranges)
    // once their flags are up
    const uint32_t ranges[3][2] = {{s->rank, s->rank + 1}, {s->rank + 1, s->world}, {0, s->rank}};
    for (int k = 0; k < 3; k++) {
        const uint32_t r0 = ranges[k][0], r1 = ranges[k][1];
        if (r0 >= r1) continue;
        const uint32_t c0 = r0 * s->cpr < s->c ? r0 * s->cpr : s->c, c1 = r1 * s->cpr < s->c ? r1 * s->cpr : s->c;
        if (k) VX_CHECK(wait_flags(s, 0, r0, r1 - r0));
        if (c0 >= c1) continue;
        VX_CHECK(lde_batch(ctx, k ? s->base + (size_t)c0 * n : mine.p, b->lde.p + (size_t)c0 * N_loc, c1 - c0, b->log_n,
                           b->rate_bits, b->blk_first, b->blk_count, b->fold_bits, b->fold_index));
    }
    VX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[0], 0));        // my own block has landed (and `mine` is free)
    VX_CUDA(cudaMemcpyAsync(b->coeffs.p, s->base, b->coeffs.bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    EV(ctx, VX_EV_LDE);
    VX_CHECK(merkle_build_device(ctx, b->lde.p, true, N_loc, N_loc, s->c, b->cap_height_loc(), b->digests.p, b->cap.p,
                                 ctx->ev[VX_EV_LEAF]));
    EV(ctx, VX_EV_TREE);
    VX_CHECK(push_block(s, ctx->stream, b->cap.p, caps_loc * 4, s->cap_off + (uint64_t)s->rank * caps_loc * 4, 1));
    VX_CHECK(wait_flags(s, 1, 0, s->world));
    if (cap_all_out)
        VX_CHECK(copy_out(ctx, cap_all_out, s->base + s->cap_off, ((size_t)4 << s->cap_height) * sizeof(u64)));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*s->err_host) {
        vx_set_error("vx_shard_commit_from_values: timed out waiting for rank %d (peer missing or failed); the group's flags "
                     "are no longer consistent: free it and create a new one", *s->err_host - 1);
        *s->err_host = 0;
        s->poisoned = true;
        return VX_ECUDA;
    }
    return VX_OK;
}

// values_local: this rank's cols_per_rank x n slice of the value columns (columns rank*cpr .. , zero rows beyond c),
// host or device.  cap_all_out: 2^cap_height x 4 (host or device, may be NULL): the cap of the WHOLE commitment.
// All ranks of the group must make this call concurrently (it waits for every peer's contribution).
extern "C" int32_t vx_shard_commit_from_values(vx_shard_group* s, const uint64_t* values_local, uint64_t* cap_all_out,
                                               vx_batch** out) {
    VX_REQUIRE(s && values_local && out, "vx_shard_commit_from_values: NULL argument");
    *out = nullptr;
    VX_REQUIRE(s->connected, "vx_shard_commit_from_values: group is not connected to its peers");
    VX_REQUIRE(!s->poisoned, "vx_shard_commit_from_values: an earlier commit on this group timed out; free it and create a new one");
    CtxGuard g(s->ctx);
    vx_batch* b = new (std::nothrow) vx_batch();
    if (!b) return VX_ENOMEM;
    const uint32_t sbits = ilog2(s->world);
    s->ctx->last_commit_lane.store(0);
    b->ctx = s->ctx; b->c = s->c; b->log_n = s->log_n; b->rate_bits = s->rate_bits; b->cap_height = s->cap_height;
    b->set_shard(s->rank, sbits);
    s->epoch++;
    int32_t r = shard_commit_run(s, b, (const u64*)values_local, (u64*)cap_all_out);
    if (r != VX_OK) {
        cudaStreamSynchronize(s->ctx->stream);
        delete b;
        return r;
    }
    *out = b;
    return VX_OK;
}
