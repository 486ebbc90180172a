// FRI batching, fold-and-commit layers, proof-of-work grinding.
//
// Replaces plonky2 v0.2.0 fri/oracle.rs `PolynomialBatch::prove_openings` (alpha-batching of ~255
// polynomials per opening point, division by X - point, LDE over the extension field) and
// fri/prover.rs `fri_committed_trees`, `fri_proof_of_work`, `fri_prover_query_rounds`; reference call
// site contracts/lib/succinctx/plonky2x/core/src/backend/circuit/build.rs:69-75 (via prove()).
// The proof layout the outputs must fill is documented in-tree at
// .../frontend/recursion/fri/proof.rs:19-42.
//
// Extension polynomials live on the device limb-major ([2][len]) so the base-field NTT kernels serve
// both limbs; the host sees them interleaved (limb0, limb1) like plonky2's QuadraticExtension.
#include "common.cuh"
#include "poseidon.cuh"

struct FriLayer {
    uint64_t leaves_n = 0;
    uint32_t width = 0, cap_height = 0;
    DevBuf leaves, digests, cap;
};

struct vx_fri {
    vx_ctx* ctx;
    uint32_t log_len;        // current length (coefficients padded to the LDE size)
    uint32_t rate_bits;
    uint64_t shift;
    DevBuf coeffs;           // [2][len]
    DevBuf values;           // [2][len], bit-reversed order
    uint32_t pending_arity = 0;
    std::vector<FriLayer*> layers;
};

// comp[m] = sum_i alpha^i * col_i[m]  (Horner from the last polynomial)
__global__ void fri_reduce_kernel(const u64* const* __restrict__ cols, uint32_t ncols, uint64_t n, gl2 alpha,
                                  u64* __restrict__ out /* [2][n] */) {
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    gl2 acc = gl2_make(0, 0);
    for (uint32_t i = ncols; i > 0; i--) acc = gl2_add_base(gl2_mul(acc, alpha), cols[i - 1][m]);
    out[m] = gl_canon(acc.a);
    out[n + m] = gl_canon(acc.b);
}

// q = (comp - comp(z)) / (X - z), i.e. q_m = c_{m+1} + z q_{m+1} (c_n = q_n = 0);  final = final * shift + q.
// One block: thread t owns a contiguous run of m.  The carry into a run is an affine function of the carry into the next
// one (carry_t = l_t + z^len_t carry_{t+1}); the 1024 runs are stitched in two levels -- lane 0 of every warp walks its 32
// runs (recording, per run, the carry for a zero warp carry-in and the multiplier of the true one), thread 0 walks the 32
// warps -- 64 dependent steps instead of 1024.
__global__ void __launch_bounds__(1024) fri_divide_accumulate_kernel(const u64* __restrict__ comp, uint64_t n, gl2 z,
                                                                    gl2 shift, u64* __restrict__ fin) {
    __shared__ u64 ca[1024], cb[1024];          // run results l_t, then the zero-carry-in carries
    __shared__ u64 pa[1024], pb[1024];          // multiplier of the warp's true carry-in for run t
    __shared__ u64 wa[32], wb[32], ma[32], mb[32], xa[32], xb[32];
    const uint64_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint64_t lo = min(n, (uint64_t)threadIdx.x * per), hi = min(n, lo + per);
    auto coef = [&](uint64_t m) { return (m < n) ? gl2_make(comp[m], comp[n + m]) : gl2_make(0, 0); };
    gl2 q = gl2_make(0, 0);
    for (uint64_t m = hi; m > lo; m--) q = gl2_add(gl2_mul(q, z), coef(m));      // q_{m-1} with zero carry-in
    ca[threadIdx.x] = q.a; cb[threadIdx.x] = q.b;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        const gl2 zp = gl2_pow(z, per);
        gl2 carry = gl2_make(0, 0), mult = gl2_make(1, 0);
        for (int t = (int)(32 * wid + 31); t >= (int)(32 * wid); t--) {
            const uint64_t tlo = min(n, (uint64_t)t * per), thi = min(n, tlo + per);
            const gl2 l = gl2_make(ca[t], cb[t]);
            ca[t] = carry.a; cb[t] = carry.b;                                     // carry into run t for a zero warp carry-in
            pa[t] = mult.a; pb[t] = mult.b;                                       // ... plus mult * (true warp carry-in)
            if (thi > tlo) {
                const gl2 zl = (thi - tlo == per) ? zp : gl2_pow(z, thi - tlo);
                carry = gl2_add(l, gl2_mul(zl, carry));
                mult = gl2_mul(zl, mult);
            }
        }
        wa[wid] = carry.a; wb[wid] = carry.b; ma[wid] = mult.a; mb[wid] = mult.b; // the warp as one affine step
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        gl2 x = gl2_make(0, 0);
        for (int w = 31; w >= 0; w--) {
            xa[w] = x.a; xb[w] = x.b;                                             // true carry into warp w
            x = gl2_add(gl2_make(wa[w], wb[w]), gl2_mul(gl2_make(ma[w], mb[w]), x));
        }
    }
    __syncthreads();
    q = gl2_add(gl2_make(ca[threadIdx.x], cb[threadIdx.x]),
                gl2_mul(gl2_make(pa[threadIdx.x], pb[threadIdx.x]), gl2_make(xa[wid], xb[wid])));
    for (uint64_t m = hi; m > lo; m--) {
        q = gl2_add(gl2_mul(q, z), coef(m));                                      // q_{m-1}
        gl2 f = gl2_make(fin[m - 1], fin[n + m - 1]);
        gl2 o = gl2_add(gl2_mul(f, shift), q);
        fin[m - 1] = gl_canon(o.a); fin[n + m - 1] = gl_canon(o.b);
    }
}

// leaves[ci][2 t + l] = values[l][ci * arity + t]
__global__ void fri_leaves_kernel(const u64* __restrict__ values, uint64_t len, uint32_t arity_bits, u64* __restrict__ leaves) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;     // element index in [0, 2 len)
    if (e >= 2 * len) return;
    uint64_t pos = e >> 1, l = e & 1;
    leaves[e] = values[l * len + pos];                                 // row-major rows of 2*arity are just the interleaving
    (void)arity_bits;
}

// coeffs'[m] = sum_{i < arity} beta^i coeffs[arity m + i]
__global__ void fri_fold_kernel(const u64* __restrict__ in, uint64_t len, uint32_t arity_bits, gl2 beta, u64* __restrict__ out) {
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t nl = len >> arity_bits;
    if (m >= nl) return;
    uint32_t ar = 1u << arity_bits;
    gl2 acc = gl2_make(0, 0);
    for (uint32_t i = ar; i > 0; i--) {
        uint64_t idx = (m << arity_bits) + i - 1;
        acc = gl2_add(gl2_mul(acc, beta), gl2_make(in[idx], in[len + idx]));
    }
    out[m] = gl_canon(acc.a);
    out[nl + m] = gl_canon(acc.b);
}

__global__ void interleave_kernel(const u64* __restrict__ in, uint64_t len, uint64_t count, u64* __restrict__ out) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * count) return;
    out[e] = in[(e & 1) * len + (e >> 1)];
}

static int32_t fri_recompute_values(vx_ctx* ctx, vx_fri* f) {
    uint64_t len = 1ULL << f->log_len;
    VX_CHECK(f->values.alloc(2 * len * 8, ctx->stream));
    VX_CUDA(cudaMemcpyAsync(f->values.p, f->coeffs.p, 2 * len * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return coset_ntt_bitrev_inplace(ctx, f->values.p, 2, f->log_len, f->shift);
}

extern "C" void vx_fri_free(vx_fri* f) {
    if (!f) return;
    {
        LaneGuard g(f->ctx);
        f->coeffs.release(); f->values.release();
        for (FriLayer* l : f->layers) { l->leaves.release(); l->digests.release(); l->cap.release(); delete l; }
    }
    delete f;
}

extern "C" int32_t vx_fri_begin(vx_ctx* ctx, vx_batch* const* oracles, uint32_t num_oracles, const vx_fri_batch* batches,
                                uint32_t num_batches, const uint64_t alpha[2], vx_fri** out) {
    VX_REQUIRE(ctx && oracles && batches && alpha && out && num_oracles >= 1 && num_batches >= 1, "vx_fri_begin: bad argument");
    *out = nullptr;
    const uint32_t log_n = oracles[0]->log_n, rate_bits = oracles[0]->rate_bits;
    for (uint32_t o = 0; o < num_oracles; o++)
        VX_REQUIRE(oracles[o] && oracles[o]->log_n == log_n && oracles[o]->rate_bits == rate_bits &&
                   oracles[o]->ctx == ctx, "vx_fri_begin: oracles must share degree, rate and context");
    VX_LANE(ctx);
    const uint64_t n = 1ULL << log_n, N = n << rate_bits;
    vx_fri* f = new (std::nothrow) vx_fri();
    if (!f) return VX_ENOMEM;
    f->ctx = ctx->root; f->log_len = log_n + rate_bits; f->rate_bits = rate_bits; f->shift = GL_GENERATOR;
    auto fail = [&](int32_t r) { cudaStreamSynchronize(ctx->stream); f->coeffs.release(); f->values.release(); delete f; return r; };
    DevBuf fin, comp, dcols;
    int32_t r = fin.alloc(2 * n * 8, ctx->stream);
    if (r == VX_OK) r = comp.alloc(2 * n * 8, ctx->stream);
    if (r != VX_OK) return fail(r);
    if (cudaMemsetAsync(fin.p, 0, 2 * n * 8, ctx->stream) != cudaSuccess) return fail(VX_ECUDA);
    gl2 al = gl2_make(alpha[0] % GL_P, alpha[1] % GL_P);
    for (uint32_t b = 0; b < num_batches; b++) {
        std::vector<const u64*> cols;
        for (uint32_t k = 0; k < batches[b].num_ranges; k++) {
            const vx_fri_range& rg = batches[b].ranges[k];
            if (rg.oracle >= num_oracles || rg.first + rg.count > oracles[rg.oracle]->c) {
                vx_set_error("vx_fri_begin: range %u of batch %u out of bounds", k, b);
                return fail(VX_EINVAL);
            }
            for (uint32_t j = 0; j < rg.count; j++) cols.push_back(oracles[rg.oracle]->coeffs.p + (uint64_t)(rg.first + j) * n);
        }
        if (cols.empty()) { vx_set_error("vx_fri_begin: empty batch %u", b); return fail(VX_EINVAL); }
        r = dcols.alloc(cols.size() * sizeof(u64*), ctx->stream);
        if (r != VX_OK) return fail(r);
        if (cudaMemcpyAsync(dcols.p, cols.data(), cols.size() * sizeof(u64*), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(VX_ECUDA);
        fri_reduce_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const u64* const*)dcols.p, (uint32_t)cols.size(), n, al, comp.p);
        // shift = alpha^(number of polynomials in this batch)
        u64 sa = 1, sb = 0;                     // host pow in the extension
        {
            auto mul2 = [](u64 a0, u64 a1, u64 b0, u64 b1, u64& c0, u64& c1) {
                u64 t0 = gl_mul_slow(a0, b0), t1 = gl_mul_slow(gl_mul_slow(a1, b1), 7);
                u64 s = t0 + t1; if (s < t0 || s >= GL_P) s -= GL_P;
                u64 u0 = gl_mul_slow(a0, b1), u1 = gl_mul_slow(a1, b0);
                u64 v = u0 + u1; if (v < u0 || v >= GL_P) v -= GL_P;
                c0 = s; c1 = v;
            };
            u64 ba = al.a, bb = al.b;
            for (size_t e = cols.size(); e; e >>= 1) {
                if (e & 1) mul2(sa, sb, ba, bb, sa, sb);
                mul2(ba, bb, ba, bb, ba, bb);
            }
        }
        gl2 z = gl2_make(batches[b].point[0] % GL_P, batches[b].point[1] % GL_P);
        fri_divide_accumulate_kernel<<<1, 1024, 0, ctx->stream>>>(comp.p, n, z, gl2_make(sa, sb), fin.p);
        VX_LAUNCH_COUNT(ctx, 2);
    }
    if (cudaGetLastError() != cudaSuccess) { vx_set_error("vx_fri_begin: kernel launch failed"); return fail(VX_ECUDA); }
    // pad to the LDE size: coeffs [2][N]
    r = f->coeffs.alloc(2 * N * 8, ctx->stream);
    if (r != VX_OK) return fail(r);
    cudaMemsetAsync(f->coeffs.p, 0, 2 * N * 8, ctx->stream);
    cudaMemcpyAsync(f->coeffs.p, fin.p, n * 8, cudaMemcpyDeviceToDevice, ctx->stream);
    cudaMemcpyAsync(f->coeffs.p + N, fin.p + n, n * 8, cudaMemcpyDeviceToDevice, ctx->stream);
    r = fri_recompute_values(ctx, f);
    if (r != VX_OK) return fail(r);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { vx_set_error("vx_fri_begin: sync failed"); return fail(VX_ECUDA); }
    *out = f;
    return VX_OK;
}

extern "C" int32_t vx_fri_commit_layer(vx_fri* f, uint32_t arity_bits, uint32_t cap_height, uint64_t* cap_out) {
    VX_REQUIRE(f && cap_out && arity_bits >= 1 && arity_bits < f->log_len, "vx_fri_commit_layer: bad argument");
    VX_REQUIRE(f->pending_arity == 0, "vx_fri_commit_layer: fold the previous layer first");
    vx_ctx* ctx = f->ctx;
    VX_LANE(ctx);
    const uint64_t len = 1ULL << f->log_len, rows = len >> arity_bits;
    VX_REQUIRE(cap_height <= f->log_len - arity_bits, "vx_fri_commit_layer: cap_height too large for this layer");
    FriLayer* L = new (std::nothrow) FriLayer();
    if (!L) return VX_ENOMEM;
    L->leaves_n = rows; L->width = 2u << arity_bits; L->cap_height = cap_height;
    int32_t r = L->leaves.alloc(2 * len * 8, ctx->stream);
    if (r == VX_OK) r = L->digests.alloc((size_t)2 * (rows - (1ULL << cap_height)) * 32, ctx->stream);
    if (r == VX_OK) r = L->cap.alloc((size_t)(1ULL << cap_height) * 32, ctx->stream);
    if (r == VX_OK) {
        fri_leaves_kernel<<<(unsigned)((2 * len + 255) / 256), 256, 0, ctx->stream>>>(f->values.p, len, arity_bits, L->leaves.p);
        VX_LAUNCH_COUNT(ctx, 1);
        r = merkle_build_device(ctx, L->leaves.p, false, 0, rows, L->width, cap_height, L->digests.p, L->cap.p);
    }
    if (r == VX_OK && cudaMemcpyAsync(cap_out, L->cap.p, L->cap.bytes, cudaMemcpyDefault, ctx->stream) != cudaSuccess) r = VX_ECUDA;
    if (r == VX_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) r = VX_ECUDA;
    if (r != VX_OK) { L->leaves.release(); L->digests.release(); L->cap.release(); delete L; return r; }
    f->layers.push_back(L);
    f->pending_arity = arity_bits;
    return VX_OK;
}

extern "C" int32_t vx_fri_fold(vx_fri* f, const uint64_t beta[2]) {
    VX_REQUIRE(f && beta, "vx_fri_fold: NULL argument");
    VX_REQUIRE(f->pending_arity != 0, "vx_fri_fold: commit the layer first");
    vx_ctx* ctx = f->ctx;
    VX_LANE(ctx);
    const uint32_t ab = f->pending_arity;
    const uint64_t len = 1ULL << f->log_len, nl = len >> ab;
    DevBuf nc;
    VX_CHECK(nc.alloc(2 * nl * 8, ctx->stream));
    fri_fold_kernel<<<(unsigned)((nl + 127) / 128), 128, 0, ctx->stream>>>(f->coeffs.p, len, ab,
                                                                         gl2_make(beta[0] % GL_P, beta[1] % GL_P), nc.p);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    VX_CHECK(f->coeffs.alloc(2 * nl * 8, ctx->stream));
    VX_CUDA(cudaMemcpyAsync(f->coeffs.p, nc.p, 2 * nl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    f->log_len -= ab;
    f->shift = gl_pow_host(f->shift, 1ULL << ab);
    f->pending_arity = 0;
    VX_CHECK(fri_recompute_values(ctx, f));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

extern "C" int32_t vx_fri_final_poly(vx_fri* f, uint64_t* out, uint32_t* len_out) {
    VX_REQUIRE(f && out && len_out, "vx_fri_final_poly: NULL argument");
    vx_ctx* ctx = f->ctx;
    VX_LANE(ctx);
    const uint64_t len = 1ULL << f->log_len, keep = len >> f->rate_bits;
    DevBuf d;
    VX_CHECK(d.alloc(2 * keep * 8, ctx->stream));
    interleave_kernel<<<(unsigned)((2 * keep + 255) / 256), 256, 0, ctx->stream>>>(f->coeffs.p, len, keep, d.p);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(out, d.p, d.bytes, cudaMemcpyDefault, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    *len_out = (uint32_t)keep;
    return VX_OK;
}

extern "C" int32_t vx_fri_query(vx_fri* f, uint32_t layer, const uint64_t* idx, uint32_t k, uint64_t* rows_out,
                                uint64_t* paths_out) {
    VX_REQUIRE(f && layer < f->layers.size() && (k == 0 || (idx && rows_out && paths_out)), "vx_fri_query: bad argument");
    if (k == 0) return VX_OK;
    vx_ctx* ctx = f->ctx;
    VX_LANE(ctx);
    FriLayer* L = f->layers[layer];
    for (uint32_t i = 0; i < k; i++)
        VX_REQUIRE(idx[i] < L->leaves_n, "vx_fri_query: leaf index %llu out of range", (unsigned long long)idx[i]);
    uint32_t depth = ilog2(L->leaves_n) - L->cap_height;
    DevBuf di, dr, dp;
    VX_CHECK(di.alloc(k * 8, ctx->stream));
    VX_CHECK(dr.alloc((size_t)k * L->width * 8, ctx->stream));
    VX_CHECK(dp.alloc((size_t)k * (depth ? depth : 1) * 32, ctx->stream));
    VX_CUDA(cudaMemcpyAsync(di.p, idx, k * 8, cudaMemcpyDefault, ctx->stream));
    VX_CHECK(gather_rows_device(ctx, L->leaves.p, false, 0, L->width, di.p, k, dr.p));
    VX_CHECK(merkle_paths_device(ctx, L->digests.p, L->leaves_n, L->cap_height, di.p, k, dp.p));
    VX_CUDA(cudaMemcpyAsync(rows_out, dr.p, (size_t)k * L->width * 8, cudaMemcpyDefault, ctx->stream));
    if (depth) VX_CUDA(cudaMemcpyAsync(paths_out, dp.p, (size_t)k * depth * 32, cudaMemcpyDefault, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

// ------------------------------------------------------------------------------------------------ proof of work
__global__ void __launch_bounds__(POSEIDON_BLOCK) pow_kernel(const u64* __restrict__ state, uint32_t pos, uint32_t min_zeros,
                                                           uint64_t base, unsigned long long* __restrict__ best) {
    __shared__ u64 scratch[12 * POSEIDON_BLOCK];
    uint64_t cand = base + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cand >= GL_P) return;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = state[i];
#pragma unroll
    for (int i = 0; i < 12; i++)
        if ((uint32_t)i == pos) s[i] = cand;
    poseidon_permute(s, scratch + threadIdx.x);
    u64 resp = gl_canon(s[7]);
    if ((resp >> (64 - min_zeros)) == 0) atomicMin(best, (unsigned long long)cand);
}

int32_t fri_module_init(vx_ctx* ctx) {
    u64 rc[360];
    poseidon_round_constants_host(rc);
    PoseidonTables t;
    if (!poseidon_derive_tables(rc, &t)) { vx_set_error("poseidon: table derivation failed"); return VX_ECUDA; }
    VX_CUDA(poseidon_upload_constants(t, ctx->stream));
    return VX_OK;
}

extern "C" int32_t vx_pow_grind(vx_ctx* ctx, const uint64_t state[12], uint32_t pos, uint32_t min_zeros, uint64_t* witness_out) {
    VX_REQUIRE(ctx && state && witness_out && pos < 8 && min_zeros >= 1 && min_zeros <= 40, "vx_pow_grind: bad argument");
    VX_LANE(ctx);
    DevBuf ds, db;
    VX_CHECK(ds.alloc(12 * 8, ctx->stream));
    VX_CHECK(db.alloc(8, ctx->stream));
    u64 st[12];
    for (int i = 0; i < 12; i++) st[i] = state[i] % GL_P;
    VX_CUDA(cudaMemcpyAsync(ds.p, st, sizeof st, cudaMemcpyHostToDevice, ctx->stream));
    unsigned long long best = ~0ULL;
    VX_CUDA(cudaMemcpyAsync(db.p, &best, 8, cudaMemcpyHostToDevice, ctx->stream));
    // candidates per launch: expected work is 2^min_zeros permutations, so the first launch tries twice that (86 % hit
    // rate, ~0.14 ms for 16 bits) and every miss doubles the batch up to 2^22
    uint64_t batch = 2ULL << min_zeros;
    if (batch < (1ULL << 14)) batch = 1ULL << 14;
    if (batch > (1ULL << 22)) batch = 1ULL << 22;
    for (uint64_t base = 0; base < GL_P; base += batch, batch = batch < (1ULL << 22) ? batch * 2 : batch) {
        pow_kernel<<<(unsigned)(batch / POSEIDON_BLOCK), POSEIDON_BLOCK, 0, ctx->stream>>>(ds.p, pos, min_zeros, base,
                                                                                        (unsigned long long*)db.p);
        VX_LAUNCH_COUNT(ctx, 1);
        VX_CUDA(cudaMemcpyAsync(&best, db.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        VX_CUDA(cudaStreamSynchronize(ctx->stream));
        if (best != ~0ULL) { *witness_out = best; return VX_OK; }   // batches are scanned in order -> smallest witness
    }
    vx_set_error("vx_pow_grind: no witness found");
    return VX_EINVAL;
}
