// Constraint evaluation over the LDE domain, Z / partial products, and opening evaluation.
//
// Replaces plonky2 v0.2.0 plonk/prover.rs `compute_quotient_polys`, plonk/vanishing_poly.rs
// `eval_vanishing_poly_base_batch`, every registered `Gate::eval_unfiltered_base_batch`
// (contracts/lib/succinctx/plonky2x/core/src/backend/circuit/serialization/gates.rs:85-107), the
// permutation-argument part of `prove_with_partition_witness`
// (.../backend/circuit/build.rs:69-75 is the reference call site) and `OpeningSet::new`.
//
// Layout: the three committed batches keep their LDE column-major in leaf (bit-reversed) order, so
// thread j evaluates LDE point bitrev(j) and every column load of a warp is 256 contiguous bytes.
// Gate constraints run as a register bytecode (include/vectorx_b200.h) whose register file is a
// [VX_PROGRAM_REGS][blockDim] shared-memory array; instruction words are warp-uniform loads.
// Constraint terms are alpha-reduced with lazy 128-bit dot products (one reduction per challenge).
#include "common.cuh"
#include "quotient_core.cuh" // with this TU's copy of the Poseidon tables: the interpreter's super-instructions run native layers

#define QCHUNK 1024          // program words staged in shared memory at a time; no instruction straddles a chunk

int32_t prover_module_init(vx_ctx* ctx) {
    u64 rc[360];
    poseidon_round_constants_host(rc);
    PoseidonTables t;
    if (!poseidon_derive_tables(rc, &t)) { vx_set_error("poseidon: table derivation failed"); return VX_ECUDA; }
    VX_CUDA(poseidon_upload_constants(t, ctx->stream));
    return VX_OK;
}

GL_D u64 lds_u64(uint32_t addr) {
    u64 v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
GL_D u64 ldg_u64(const u64* p) {
    u64 v;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
// hot-operation flags in the top bits of a device instruction word's immediate (bits 0..25 stay the immediate)
#define QF_LOADW 0x80000000u
#define QF_EMIT 0x40000000u
#define QF_MADK 0x20000000u
#define QF_RANGE4 0x10000000u
#define QF_SUB 0x08000000u
#define QF_MUL 0x04000000u
#define QF_WAIT 0x02000000u        // inserted by the host after every run of column loads: wait for the asynchronous copies
#define QIMM_MASK 0x01ffffffu
#define VX_OP_WAIT_INTERNAL 0xfe
// column load: asynchronous 8-byte global -> shared copy straight into the register file (no stall until the run's WAIT)
GL_D void cp_async_u64(uint32_t saddr, const u64* g, u64 pol) {
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(saddr), "l"(g), "l"(pol) : "memory");
}
GL_D void cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
GL_D void sts_u64(uint32_t addr, u64 v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }

// 12 register numbers from two operand words
#define QREGS12(w0, w1, k) (((k) < 8 ? (uint32_t)((w0) >> (8 * (k))) : (uint32_t)((w1) >> (8 * ((k) - 8)))) & 0xffu)

#ifndef QUOT_MINB
#define QUOT_MINB 5            // 96 registers: 4.97 ms at 2^16 rows; 114 registers (no bound) 5.26 ms, 80 registers 5.02 ms
#endif
__global__ void __launch_bounds__(QBLOCK, QUOT_MINB) quotient_kernel(const QuotParams p) {
    extern __shared__ u64 qsmem[];
    u64* prog_s = qsmem;                                   // [QCHUNK]
    u64* scratch = qsmem + QCHUNK + threadIdx.x;           // [12][QBLOCK] column of this thread (dense layer)
    u64* R = qsmem + QCHUNK + 12 * QBLOCK + threadIdx.x;   // [num_regs][QBLOCK] register file column
    GlAcc2 tot[2];                      // indexed statically everywhere: a run-time index would park both in local memory
    const QuotPoint q = quot_prologue(p, tot);
    const uint64_t j = q.j, N = p.N;
    const uint32_t nch = p.num_challenges;
    const u64* apow0 = p.apow;
    const u64* apow1 = p.apow + p.num_terms;

    // ---- gate constraints: bytecode interpreter.  The program is the same for every thread: the block stages it through
    // shared memory QCHUNK words at a time, so an instruction fetch is a broadcast LDS instead of a dependent global load.
    // Register file and program are addressed with 32-bit shared-window addresses (ld/st.shared): an operand access is
    // one shift-add + one LDS, and the per-point column pointers are pinned so that nothing is recomputed per
    // instruction (the first form of this loop spent 41 of its ~87 instructions per bytecode operation on decode).
    GlAcc2 h[2];
    uint32_t cidx = 0;
    bool done = false;
    uint32_t prog_sa = (uint32_t)__cvta_generic_to_shared(prog_s);
    uint32_t reg_sa = (uint32_t)__cvta_generic_to_shared(R);
    const u64* wires_j = p.wires + j;
    const u64* cs_j = p.cs + j;
    uint64_t Nq = N;
    const u64 pol_keep = quot_policy_keep();       // gate columns are read again by other gates: keep them in L2
    // opaque to the optimiser: held in registers instead of being recomputed from special registers per instruction
    asm volatile("" : "+l"(wires_j), "+l"(cs_j), "+r"(prog_sa), "+r"(reg_sa), "+l"(Nq));
#define RLD(i) lds_u64(reg_sa + ((i) << 10))
#define RST(i, v) sts_u64(reg_sa + ((i) << 10), (v))
    static_assert(QBLOCK * sizeof(u64) == 1024, "register-file row pitch is 1 KB");
    for (uint32_t base = 0; base < p.program_len && !done; base += QCHUNK) {
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < QCHUNK; t += QBLOCK) prog_s[t] = __ldg(p.program + base + t);
        __syncthreads();
        uint32_t pa = prog_sa;                               // shared address of the next instruction word
        const uint32_t pend = prog_sa + QCHUNK * 8;
        while (pa < pend) {
            const u64 ins = lds_u64(pa);
            pa += 8;
            // device word (re-laid by vx_quotient): the most frequent operations are flagged in the top bits of the
            // immediate, tested one bit at a time in order of frequency (a switch compiles to a balanced tree instead)
            const uint32_t lo = (uint32_t)ins, hi = (uint32_t)(ins >> 32), imm = hi & QIMM_MASK;
            const uint32_t dst = (lo >> 8) & 0xff, ra = (lo >> 16) & 0xff, rb = lo >> 24;
            if (hi & QF_LOADW) { cp_async_u64(reg_sa + (dst << 10), wires_j + (uint64_t)imm * Nq, pol_keep); continue; }
            if (hi & QF_EMIT) {
                const u64 v = RLD(ra);
                gl_acc2_mad(h[0], v, __ldg(apow0 + cidx));
                if (nch > 1) gl_acc2_mad(h[1], v, __ldg(apow1 + cidx));
                cidx++;
                continue;
            }
            if (hi & QF_MADK) { const u64 k = lds_u64(pa); pa += 8; RST(dst, gl_mul_add_cc(RLD(ra), k, RLD(rb))); continue; }
            if (hi & QF_RANGE4) {
                // a (a-1)(a-2)(a-3) = y (y + 2) with y = a (a - 3): two multiplications instead of three
                const u64 a = RLD(ra);
                const u64 y = gl_mul_cc(a, gl_sub(a, 3));
                RST(dst, gl_mul_cc(y, gl_add(y, 2)));
                continue;
            }
            if (hi & QF_WAIT) { cp_async_wait(); continue; }
            if (hi & QF_SUB) { RST(dst, gl_sub(RLD(ra), RLD(rb))); continue; }
            if (hi & QF_MUL) { RST(dst, gl_mul_cc(RLD(ra), RLD(rb))); continue; }
            const uint32_t op = lo & 0xff;
            if (op == VX_OP_END) { done = true; break; }
            switch (op) {
                case VX_OP_SBOX7: RST(dst, gl_pow7_cc(RLD(ra))); break;
                case VX_OP_LOADC: cp_async_u64(reg_sa + (dst << 10), cs_j + (uint64_t)imm * Nq, pol_keep); break;
                case VX_OP_LOADPI: RST(dst, p.pi_hash[imm & 3]); break;
                case VX_OP_LOADK: RST(dst, lds_u64(pa)); pa += 8; break;
                case VX_OP_ADD: RST(dst, gl_add(RLD(ra), RLD(rb))); break;
                case VX_OP_ADDK: RST(dst, gl_add(RLD(ra), lds_u64(pa))); pa += 8; break;
                case VX_OP_MULK: RST(dst, gl_mul_cc(RLD(ra), lds_u64(pa))); pa += 8; break;
                case VX_OP_RSUBK: RST(dst, gl_sub(lds_u64(pa), RLD(ra))); pa += 8; break;
                case VX_OP_SUBK: RST(dst, gl_sub(RLD(ra), lds_u64(pa))); pa += 8; break;
                case VX_OP_MDS12K:
                case VX_OP_DENSE12:
                case VX_OP_PARTIAL12: {
                    const u64 s0 = lds_u64(pa), s1 = lds_u64(pa + 8), d0 = lds_u64(pa + 16), d1 = lds_u64(pa + 24);
                    pa += 32;
                    u64 st[12];
#pragma unroll
                    for (int k = 0; k < 12; k++) st[k] = RLD(QREGS12(s0, s1, k));
                    if (op == VX_OP_MDS12K) {
                        poseidon_mds_add_freq(st, c_pos.rc22 + 36 * imm);
                    } else if (op == VX_OP_DENSE12) {
                        poseidon_dense_layer(st, scratch);
                    } else {
                        quot_partial12(st, imm);
                    }
#pragma unroll
                    for (int k = 0; k < 12; k++) RST(QREGS12(d0, d1, k), st[k]);
                    break;
                }
                case VX_OP_BEGINGATE:
                    gl_acc2_init(h[0], 0); gl_acc2_init(h[1], 0);
                    cidx = p.num_perm_terms;
                    break;
                case VX_OP_ENDGATE: {
                    const u64 f = (ra == 255) ? 1 : RLD(ra);
                    gl_acc2_mad(tot[0], f, gl_acc2_reduce(h[0]));
                    if (nch > 1) gl_acc2_mad(tot[1], f, gl_acc2_reduce(h[1]));
                    break;
                }
                default: break;                     // VX_OP_NOP
            }
        }
    }
#undef RLD
#undef RST
    quot_epilogue(p, q, tot);
}

// out[col][bitrev(p)] = in[col][p]
__global__ void bitrev_permute_kernel(const u64* __restrict__ in, u64* __restrict__ out, uint32_t bits) {
    uint64_t pidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pidx >= (1ULL << bits)) return;
    uint64_t base = (uint64_t)blockIdx.y << bits;
    out[base + bitrev_u64(pidx, bits)] = in[base + pidx];
}

extern "C" int32_t vx_quotient(vx_ctx* ctx, const vx_circuit_desc* d, vx_batch* cs, vx_batch* wires, vx_batch* zpp,
                               const uint64_t pi_hash[4], const uint64_t* betas, const uint64_t* gammas,
                               const uint64_t* alphas, uint64_t* out) {
    VX_REQUIRE(ctx && d && cs && wires && zpp && pi_hash && betas && gammas && alphas && out, "vx_quotient: NULL argument");
    VX_REQUIRE(d->num_challenges >= 1 && d->num_challenges <= 2, "vx_quotient: num_challenges %u unsupported", d->num_challenges);
    VX_REQUIRE(d->rate_bits <= 6, "vx_quotient: rate_bits %u unsupported", d->rate_bits);
    const uint32_t bits = d->degree_bits + d->rate_bits;
    const uint64_t N = 1ULL << bits, n = 1ULL << d->degree_bits;
    VX_REQUIRE(cs->N_loc() == N && wires->N_loc() == N && zpp->N_loc() == N &&
               cs->log_n == d->degree_bits && wires->log_n == d->degree_bits && zpp->log_n == d->degree_bits,
               "vx_quotient: batches must be whole (unsharded) commitments of degree 2^%u, rate %u", d->degree_bits, d->rate_bits);
    VX_REQUIRE(wires->c == d->num_wires && cs->c == d->num_constants + d->num_routed_wires &&
               zpp->c == d->num_challenges * (1 + d->num_partial_products), "vx_quotient: batch widths do not match the circuit");
    const uint32_t chunks = (d->num_routed_wires + d->max_degree - 1) / d->max_degree;
    VX_REQUIRE(chunks == d->num_partial_products + 1, "vx_quotient: num_partial_products inconsistent with max_degree");
    VX_LANE(ctx);
    const uint32_t nch = d->num_challenges;
    const uint32_t perm_terms = nch + nch * chunks;
    const uint32_t num_terms = perm_terms + d->num_gate_constraints;
    std::vector<u64> h_apow((size_t)nch * num_terms), h_betak((size_t)nch * d->num_routed_wires);
    for (uint32_t k = 0; k < nch; k++) {
        u64 acc = 1, al = alphas[k] % GL_P;
        for (uint32_t t = 0; t < num_terms; t++) { h_apow[(size_t)k * num_terms + t] = acc; acc = gl_mul_slow(acc, al); }
        for (uint32_t w = 0; w < d->num_routed_wires; w++)
            h_betak[(size_t)k * d->num_routed_wires + w] = gl_mul_slow(betas[k] % GL_P, d->k_is[w] % GL_P);
    }
    DevBuf d_apow, d_betak, d_prog, d_q, d_qnat;
    VX_CHECK(d_apow.alloc(h_apow.size() * 8, ctx->stream));
    VX_CHECK(d_betak.alloc(h_betak.size() * 8, ctx->stream));
    VX_CHECK(d_q.alloc((size_t)nch * N * 8, ctx->stream));
    VX_CHECK(d_qnat.alloc((size_t)nch * N * 8, ctx->stream));
    VX_CUDA(cudaMemcpyAsync(d_apow.p, h_apow.data(), h_apow.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(cudaMemcpyAsync(d_betak.p, h_betak.data(), h_betak.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    // the circuit's gate program compiled at load time (vx_quotient_compile), if there is one: nothing to stage then
    uint32_t jit_threads = QBLOCK;
    const void* jit = quotient_jit_lookup(ctx, d, &jit_threads);
    // re-lay the program so that no instruction straddles a QCHUNK boundary (pad with NOPs), find the register count
    std::vector<u64> prog;
    prog.reserve(d->program_len + 2 * QCHUNK);
    uint32_t max_reg = 0;
    bool loads_pending = false;
    auto oplen = [](uint32_t op) -> uint32_t {
        switch (op) {
            case VX_OP_LOADK: case VX_OP_ADDK: case VX_OP_MULK: case VX_OP_RSUBK: case VX_OP_SUBK: case VX_OP_MADK: return 2;
            case VX_OP_MDS12K: case VX_OP_DENSE12: case VX_OP_PARTIAL12: return 5;
            default: return 1;
        }
    };
    for (uint64_t pc = 0; pc < d->program_len && !jit;) {
        const u64 ins = d->program[pc];
        const uint32_t op = (uint32_t)ins & 0xff, len = oplen(op);
        VX_REQUIRE(op <= VX_OP_MADK && pc + len <= d->program_len, "vx_quotient: malformed program at word %llu",
                   (unsigned long long)pc);
        const bool is_load = op == VX_OP_LOADW || op == VX_OP_LOADC;
        if (loads_pending && !is_load) {                       // end of a run of column loads: one wait for all of them
            prog.push_back((u64)VX_OP_WAIT_INTERNAL | ((u64)QF_WAIT << 32));
            loads_pending = false;
        }
        if (is_load) loads_pending = true;
        while (prog.size() % QCHUNK + len > QCHUNK) prog.push_back(VX_OP_NOP);
        VX_REQUIRE((ins >> 32) <= QIMM_MASK, "vx_quotient: immediate of the instruction at word %llu out of range",
                   (unsigned long long)pc);
        uint32_t flag = 0;
        switch (op) {
            case VX_OP_LOADW: flag = QF_LOADW; break;
            case VX_OP_EMIT: flag = QF_EMIT; break;
            case VX_OP_MADK: flag = QF_MADK; break;
            case VX_OP_RANGE4: flag = QF_RANGE4; break;
            case VX_OP_SUB: flag = QF_SUB; break;
            case VX_OP_MUL: flag = QF_MUL; break;
            default: break;
        }
        prog.push_back(ins | ((u64)flag << 32));
        for (uint32_t k = 1; k < len; k++) prog.push_back(d->program[pc + k]);
        if (len == 5) {
            for (uint32_t w = 1; w < 5; w++)
                for (int k = 0; k < ((w & 1) ? 8 : 4); k++)
                    max_reg = std::max(max_reg, (uint32_t)((d->program[pc + w] >> (8 * k)) & 0xff));
        } else if (op != VX_OP_BEGINGATE && op != VX_OP_NOP && op != VX_OP_END) {
            const uint32_t dst = (uint32_t)(ins >> 8) & 0xff, ra = (uint32_t)(ins >> 16) & 0xff, rb = (uint32_t)(ins >> 24) & 0xff;
            max_reg = std::max(max_reg, dst);
            if (!(op == VX_OP_ENDGATE && ra == 255)) max_reg = std::max(max_reg, ra);
            max_reg = std::max(max_reg, rb);
        }
        pc += len;
    }
    VX_REQUIRE(max_reg < VX_PROGRAM_REGS, "vx_quotient: program uses register %u (limit %d)", max_reg, VX_PROGRAM_REGS);
    prog.push_back(VX_OP_END);
    while (prog.size() % QCHUNK) prog.push_back(VX_OP_END);
    if (!jit) {
        VX_CHECK(d_prog.alloc(prog.size() * 8, ctx->stream));
        VX_CUDA(cudaMemcpyAsync(d_prog.p, prog.data(), prog.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    }

    QuotParams p;
    memset(&p, 0, sizeof p);
    p.cs = cs->lde.p; p.wires = wires->lde.p; p.zpp = zpp->lde.p;
    p.N = N; p.bits = bits; p.rate_bits = d->rate_bits; p.degree_bits = d->degree_bits;
    p.num_wires = d->num_wires; p.num_routed = d->num_routed_wires; p.num_constants = d->num_constants;
    p.num_selectors = d->num_selectors; p.num_challenges = nch; p.num_pp = d->num_partial_products;
    p.max_degree = d->max_degree;
    p.program = d_prog.p; p.program_len = (uint32_t)prog.size(); p.num_regs = max_reg + 1;
    p.beta_k = d_betak.p; p.apow = d_apow.p;
    p.num_terms = num_terms; p.num_perm_terms = perm_terms;
    for (uint32_t k = 0; k < nch; k++) { p.betas[k] = betas[k] % GL_P; p.gammas[k] = gammas[k] % GL_P; }
    for (int t = 0; t < 4; t++) p.pi_hash[t] = pi_hash[t] % GL_P;
    // Z_H(g w_N^i) = g^n * (w_N^n)^i - 1 depends on i mod 2^rate_bits only
    u64 gn = gl_pow_host(GL_GENERATOR, n), wr = gl_root_of_unity_host(d->rate_bits), acc = 1;
    for (uint32_t c = 0; c < (1u << d->rate_bits); c++) {
        u64 v = gl_mul_slow(gn, acc);
        v = v ? v - 1 : GL_P - 1;
        VX_REQUIRE(v != 0, "vx_quotient: Z_H vanishes on the coset");
        p.zh[c] = v;
        p.zh_inv[c] = gl_inv_host(v);
        acc = gl_mul_slow(acc, wr);
    }
    p.n_inv = gl_inv_host(n % GL_P);
    p.out = d_q.p;
    p.tw.lo = ctx->w_lo; p.tw.hi = ctx->w_hi; p.tw.roots12 = ctx->roots12; p.tw.full12 = ctx->roots12f;
    if (jit) {
        // same arithmetic as the interpreter below, no decode
        void* args[] = {(void*)&p};
        VX_CUDA(cudaLaunchKernel(jit, dim3((unsigned)((N + jit_threads - 1) / jit_threads)), dim3(jit_threads), args, 0, ctx->stream));
    } else {
        size_t smem = ((size_t)QCHUNK + (size_t)(12 + p.num_regs) * QBLOCK) * sizeof(u64);
        VX_CUDA(cudaFuncSetAttribute(quotient_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        quotient_kernel<<<(unsigned)((N + QBLOCK - 1) / QBLOCK), QBLOCK, smem, ctx->stream>>>(p);
    }
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    // leaf order -> natural order, then coset iNTT (transpose + coset_ifft(g) upstream)
    dim3 grid((unsigned)((N + 255) / 256), nch);
    bitrev_permute_kernel<<<grid, 256, 0, ctx->stream>>>(d_q.p, d_qnat.p, bits);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CHECK(ntt_natural(ctx, d_qnat.p, d_q.p, nch, bits, true, GL_GENERATOR));
    VX_CUDA(cudaMemcpyAsync(out, d_q.p, (size_t)nch * N * 8, cudaMemcpyDefault, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

// ------------------------------------------------------------------------------------------------ Z / partial products
// phase 1: per (challenge, row): chunk quotients q_c = prod num / prod den and their row product.  The `chunks`
// denominators of a row are inverted together (Montgomery's trick: one field inversion per row instead of one per
// chunk); numerators, denominators and their running products pass through global scratch (coalesced, chunk-major).
__global__ void zpp_rows_kernel(const u64* __restrict__ wires, const u64* __restrict__ sigmas, uint64_t n,
                                uint32_t log_n, uint32_t num_routed, uint32_t max_degree, uint32_t chunks,
                                const u64* __restrict__ beta_k, u64 beta, u64 gamma, TwiddleView tw,
                                u64* __restrict__ q /* [chunks][n] */, u64* __restrict__ dtmp /* [chunks][n] */,
                                u64* __restrict__ ptmp /* [chunks][n] */, u64* __restrict__ rowprod /* [n] */) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    u64 s = tw_pow_view(tw, (u32)(r << (32 - log_n)));          // w_n^r
    u64 run = 1;
    for (uint32_t c = 0; c < chunks; c++) {
        u64 num = 1, den = 1;
        uint32_t hi = min(num_routed, (c + 1) * max_degree);
        for (uint32_t w = c * max_degree; w < hi; w++) {
            u64 wv = wires[(uint64_t)w * n + r];
            u64 a = gl_add(gl_mul_add_cc(s, __ldg(beta_k + w), wv), gamma);
            u64 b = gl_add(gl_mul_add_cc(sigmas[(uint64_t)w * n + r], beta, wv), gamma);
            num = gl_mul_cc(num, a);
            den = gl_mul_cc(den, b);
        }
        q[(uint64_t)c * n + r] = num;
        dtmp[(uint64_t)c * n + r] = den;
        ptmp[(uint64_t)c * n + r] = run;                          // product of the denominators before chunk c
        run = gl_mul_cc(run, den);
    }
    u64 inv = gl_inv(run);                                        // 1 / (den_0 ... den_{chunks-1})
    u64 rp = 1;
    for (uint32_t c = chunks; c-- > 0;) {
        const u64 inv_den = gl_mul_cc(inv, ptmp[(uint64_t)c * n + r]);
        inv = gl_mul_cc(inv, dtmp[(uint64_t)c * n + r]);
        const u64 qc = gl_canon(gl_mul_cc(q[(uint64_t)c * n + r], inv_den));
        q[(uint64_t)c * n + r] = qc;
        rp = gl_mul_cc(rp, qc);
    }
    rowprod[r] = gl_canon(rp);
}

// phase 2: exclusive prefix product over rows, one block of 1024 threads; thread t owns a contiguous run of rows and the
// 1024 run products are scanned with warp shuffles (5 + 5 dependent multiplications instead of a serial loop)
__global__ void __launch_bounds__(1024) zpp_scan_kernel(const u64* __restrict__ rowprod, uint64_t n, u64* __restrict__ z) {
    __shared__ u64 warp_tot[32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t per = (n + blockDim.x - 1) / blockDim.x;
    uint64_t lo = min(n, (uint64_t)threadIdx.x * per), hi = min(n, lo + per);
    u64 acc = 1;
    for (uint64_t r = lo; r < hi; r++) acc = gl_mul_cc(acc, rowprod[r]);
    // inclusive scan inside the warp
    u64 inc = acc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 up = __shfl_up_sync(0xffffffffu, inc, d);
        if ((int)lane >= d) inc = gl_mul_cc(inc, up);
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        u64 t = warp_tot[lane], ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 up = __shfl_up_sync(0xffffffffu, ti, d);
            if ((int)lane >= d) ti = gl_mul_cc(ti, up);
        }
        u64 ex = __shfl_up_sync(0xffffffffu, ti, 1);              // exclusive: product of the warps before this one
        warp_tot[lane] = lane ? ex : 1;
    }
    __syncthreads();
    u64 before = __shfl_up_sync(0xffffffffu, inc, 1);             // product of the lanes before this one in the warp
    if (lane == 0) before = 1;
    acc = gl_mul_cc(warp_tot[wid], before);
    for (uint64_t r = lo; r < hi; r++) { z[r] = gl_canon(acc); acc = gl_mul_cc(acc, rowprod[r]); }
}

// phase 3: partial products inside each row
__global__ void zpp_fill_kernel(const u64* __restrict__ q, const u64* __restrict__ z, uint64_t n, uint32_t chunks,
                                u64* __restrict__ pp /* [chunks-1][n] */) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    u64 acc = z[r];
    for (uint32_t c = 0; c + 1 < chunks; c++) {
        acc = gl_mul_cc(acc, q[(uint64_t)c * n + r]);
        pp[(uint64_t)c * n + r] = gl_canon(acc);
    }
}

extern "C" int32_t vx_zs_partial_products(vx_ctx* ctx, const vx_circuit_desc* d, const uint64_t* wires,
                                          const uint64_t* sigmas, const uint64_t* betas, const uint64_t* gammas,
                                          uint64_t* out) {
    VX_REQUIRE(ctx && d && wires && sigmas && betas && gammas && out, "vx_zs_partial_products: NULL argument");
    VX_LANE(ctx);
    const uint64_t n = 1ULL << d->degree_bits;
    const uint32_t nch = d->num_challenges, R = d->num_routed_wires;
    VX_REQUIRE(nch >= 1 && nch <= 4, "vx_zs_partial_products: num_challenges %u unsupported", nch);
    const uint32_t chunks = (R + d->max_degree - 1) / d->max_degree;
    VX_REQUIRE(chunks == d->num_partial_products + 1, "vx_zs_partial_products: num_partial_products inconsistent");
    // inputs / output already in device memory (the prover keeps them there) are used in place
    const bool w_dev = vx_is_device_ptr(wires), s_dev = vx_is_device_ptr(sigmas), o_dev = vx_is_device_ptr(out);
    DevBuf dw, ds, dq, dd, dp, drp, dout, dbk;
    if (!w_dev) {
        VX_CHECK(dw.alloc((size_t)R * n * 8, ctx->stream));       // only routed wires take part
        VX_CUDA(cudaMemcpyAsync(dw.p, wires, (size_t)R * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!s_dev) {
        VX_CHECK(ds.alloc((size_t)R * n * 8, ctx->stream));
        VX_CUDA(cudaMemcpyAsync(ds.p, sigmas, (size_t)R * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    const u64* pw = w_dev ? (const u64*)wires : dw.p;
    const u64* ps = s_dev ? (const u64*)sigmas : ds.p;
    VX_CHECK(dq.alloc((size_t)chunks * n * 8, ctx->stream));
    VX_CHECK(dd.alloc((size_t)chunks * n * 8, ctx->stream));
    VX_CHECK(dp.alloc((size_t)chunks * n * 8, ctx->stream));
    VX_CHECK(drp.alloc(n * 8, ctx->stream));
    if (!o_dev) VX_CHECK(dout.alloc((size_t)nch * chunks * n * 8, ctx->stream));
    u64* po = o_dev ? (u64*)out : dout.p;
    VX_CHECK(dbk.alloc((size_t)nch * R * 8, ctx->stream));
    TwiddleView tw; tw.lo = ctx->w_lo; tw.hi = ctx->w_hi; tw.roots12 = ctx->roots12; tw.full12 = ctx->roots12f;
    std::vector<u64> bk((size_t)nch * R);
    for (uint32_t k = 0; k < nch; k++)
        for (uint32_t w = 0; w < R; w++) bk[(size_t)k * R + w] = gl_mul_slow(betas[k] % GL_P, d->k_is[w] % GL_P);
    VX_CUDA(cudaMemcpyAsync(dbk.p, bk.data(), bk.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    for (uint32_t k = 0; k < nch; k++) {
        unsigned blocks = (unsigned)((n + 127) / 128);
        u64* z = po + (size_t)k * n;
        u64* pp = po + (size_t)nch * n + (size_t)k * (chunks - 1) * n;
        zpp_rows_kernel<<<blocks, 128, 0, ctx->stream>>>(pw, ps, n, d->degree_bits, R, d->max_degree, chunks,
                                                         dbk.p + (size_t)k * R, betas[k] % GL_P, gammas[k] % GL_P, tw,
                                                         dq.p, dd.p, dp.p, drp.p);
        zpp_scan_kernel<<<1, 1024, 0, ctx->stream>>>(drp.p, n, z);
        zpp_fill_kernel<<<blocks, 128, 0, ctx->stream>>>(dq.p, z, n, chunks, pp);
        VX_LAUNCH_COUNT(ctx, 3);
    }
    VX_CUDA(cudaGetLastError());
    if (!o_dev) VX_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)nch * chunks * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));                 // also keeps `bk` alive until its copy has been issued
    return VX_OK;
}

// ------------------------------------------------------------------------------------------------ openings
// One block per polynomial: thread t evaluates the coefficients m = t (mod block) by Horner in point^block, scales by
// point^t and the block sums the pieces.  out[col] = sum_m coeffs[col][m] * point^m  (extension value).
#define EVAL_BLOCK 1024
__global__ void __launch_bounds__(EVAL_BLOCK) eval_ext_kernel(const u64* __restrict__ coeffs, uint32_t log_n, gl2 point,
                                                       u64* __restrict__ out) {
    __shared__ u64 sa[EVAL_BLOCK], sb[EVAL_BLOCK];
    const uint64_t n = 1ULL << log_n;
    const u64* col = coeffs + ((uint64_t)blockIdx.x << log_n);
    // thread t takes the coefficients m = t (mod EVAL_BLOCK): Horner in point^EVAL_BLOCK over coalesced loads, then x point^t
    const gl2 step = gl2_pow(point, EVAL_BLOCK);
    gl2 acc = gl2_make(0, 0);
    const uint64_t runs = (n + EVAL_BLOCK - 1) / EVAL_BLOCK;
    for (uint64_t k = runs; k > 0; k--) {
        const uint64_t m = (k - 1) * EVAL_BLOCK + threadIdx.x;
        acc = gl2_mul(acc, step);
        if (m < n) acc = gl2_add_base(acc, col[m]);
    }
    acc = gl2_mul(acc, gl2_pow(point, threadIdx.x));
    sa[threadIdx.x] = gl_canon(acc.a); sb[threadIdx.x] = gl_canon(acc.b);
    __syncthreads();
    for (int s = EVAL_BLOCK / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            sa[threadIdx.x] = gl_canon(gl_add(sa[threadIdx.x], sa[threadIdx.x + s]));
            sb[threadIdx.x] = gl_canon(gl_add(sb[threadIdx.x], sb[threadIdx.x + s]));
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = sa[0]; out[2 * blockIdx.x + 1] = sb[0]; }
}

extern "C" int32_t vx_batch_eval_ext(vx_batch* b, const uint64_t point[2], uint64_t* out) {
    VX_REQUIRE(b && point && out, "vx_batch_eval_ext: NULL argument");
    vx_ctx* ctx = b->ctx;
    VX_LANE(ctx);
    DevBuf d;
    VX_CHECK(d.alloc((size_t)b->c * 2 * 8, ctx->stream));
    eval_ext_kernel<<<b->c, EVAL_BLOCK, 0, ctx->stream>>>(b->coeffs.p, b->log_n, gl2_make(point[0] % GL_P, point[1] % GL_P), d.p);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(out, d.p, d.bytes, cudaMemcpyDefault, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

// ------------------------------------------------------------------------------------------------ field self-test surface
__global__ void field_op_kernel(uint32_t op, const u64* __restrict__ a, const u64* __restrict__ b, const u64* __restrict__ c,
                                uint64_t n, u64* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (op) {
        case 0: out[i] = gl_canon(gl_mul_cc(a[i], b[i])); break;
        case 1: out[i] = gl_canon(gl_mul(a[i], b[i])); break;
        case 2: out[i] = gl_canon(gl_inv(a[i])); break;
        case 3: out[i] = gl_canon(gl_mul_add_cc(a[i], b[i], c[i])); break;
        case 4: out[i] = gl_canon(gl_add(a[i], b[i])); break;
        case 5: out[i] = gl_canon(gl_sub(a[i], b[i])); break;
        case 6: out[i] = gl_canon(gl_pow7_cc(a[i])); break;
        case 7: { gl2 r = gl2_canon(gl2_mul(gl2_make(a[2 * i], a[2 * i + 1]), gl2_make(b[2 * i], b[2 * i + 1])));
                  out[2 * i] = r.a; out[2 * i + 1] = r.b; break; }
        case 8: { gl2 r = gl2_canon(gl2_inv(gl2_make(a[2 * i], a[2 * i + 1]))); out[2 * i] = r.a; out[2 * i + 1] = r.b; break; }
        default: break;
    }
}

extern "C" int32_t vx_field_op(vx_ctx* ctx, uint32_t op, const uint64_t* a, const uint64_t* b, const uint64_t* c, uint64_t n,
                               uint64_t* out) {
    VX_REQUIRE(ctx && a && out && op <= 8, "vx_field_op: bad argument");
    if (n == 0) return VX_OK;
    VX_LANE(ctx);
    const uint64_t words = (op >= 7) ? 2 * n : n;
    DevBuf da, db, dc, dout;
    VX_CHECK(da.alloc(words * 8, ctx->stream)); VX_CHECK(db.alloc(words * 8, ctx->stream));
    VX_CHECK(dc.alloc(words * 8, ctx->stream)); VX_CHECK(dout.alloc(words * 8, ctx->stream));
    VX_CUDA(cudaMemcpyAsync(da.p, a, words * 8, cudaMemcpyDefault, ctx->stream));
    if (b) VX_CUDA(cudaMemcpyAsync(db.p, b, words * 8, cudaMemcpyDefault, ctx->stream));
    if (c) VX_CUDA(cudaMemcpyAsync(dc.p, c, words * 8, cudaMemcpyDefault, ctx->stream));
    field_op_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(op, da.p, db.p, dc.p, n, dout.p);
    VX_LAUNCH_COUNT(ctx, 1);
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(out, dout.p, words * 8, cudaMemcpyDefault, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}
