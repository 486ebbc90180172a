/* vectorx_b200 -- C ABI of the B200 (sm_100a) proving hot path behind VectorX's plonky2x proofs.
 *
 * This header is the drop-in boundary: each entry point is what a patched plonky2 v0.2.0
 * (`[patch."https://github.com/0xPolygonZero/plonky2.git"]`) binds through
 * `extern "C"` in place of the Rust body named in its comment.  plonky2 itself is NOT vendored in
 * the reference (Cargo.lock:4847-4850); "replaces" cites the upstream item and the reference call
 * site that reaches it.  P2X = contracts/lib/succinctx/plonky2x/core/src (under /root/reference).
 *
 * Conventions
 *  - field elements are little-endian u64 Goldilocks values (p = 2^64 - 2^32 + 1); any
 *    representative is accepted on input, outputs are always canonical (< p);
 *  - every call returns 0 on success or a negative VX_E* code; no exception crosses the boundary;
 *    vx_last_error() returns a thread-local message for the last failure;
 *  - host pointers may be pageable; pointers marked "host or device" are classified with
 *    cudaPointerGetAttributes;
 *  - the caller owns host buffers, the library owns device memory behind the opaque handles;
 *  - all entry points are thread-safe (callers are Rayon workers: P2X/backend/circuit/build.rs:127,
 *    P2X/frontend/hint/synchronous.rs:41); each call is blocking (outputs valid on return).
 *  - there is no CPU fallback: if no sm_100-class device is usable vx_ctx_create fails.
 */
#ifndef VECTORX_B200_H
#define VECTORX_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VX_OK 0
#define VX_EINVAL (-1)    /* bad argument */
#define VX_ECUDA (-2)     /* CUDA runtime error (message has the detail) */
#define VX_ENOMEM (-3)    /* device allocation failed */
#define VX_ENODEV (-4)    /* no usable GPU */
#define VX_EUNSUPPORTED (-5)

/* Per-device context: twiddle / constant tables and a pool of independent stream sets ("lanes").  Threading contract
 * (callers are arbitrary Rayon worker threads and a tokio block_in_place thread, several STARK proofs in
 * flight -- P2X/backend/circuit/build.rs:69-75,127, P2X/frontend/hint/synchronous.rs:41): every entry point is
 * re-entrant and blocking; a call takes a free lane for its duration, so up to 4 calls on one context run concurrently
 * on the device and further callers queue.  Handles (vx_batch, vx_tree, vx_fri) may be used from any thread, one call at
 * a time per handle. */
typedef struct vx_ctx vx_ctx;
typedef struct vx_batch vx_batch;   /* device-resident PolynomialBatch (coeffs + LDE + Merkle tree) */
typedef struct vx_tree vx_tree;     /* device-resident MerkleTree over row-major leaves */

/* ---- context ------------------------------------------------------------------------------ */
int32_t vx_ctx_create(int32_t device, vx_ctx** out);
/* number of sm_100-class devices a context can be created on (0 when there is none; never fails).  The Rust side sizes
 * its proof-level fan-out with it: LocalProver::batch_prove, P2X/backend/prover/local.rs:44-48, proves its independent
 * inputs in a sequential loop; with one context per device they run side by side. */
int32_t vx_device_count(void);
/* CUDA ordinals of those devices (a box may mix GPU generations, so they need not be 0..count-1): writes up to `capacity`
 * ordinals and returns how many there are. */
int32_t vx_device_list(int32_t* ordinals_out, int32_t capacity);
void vx_ctx_destroy(vx_ctx* ctx);
const char* vx_last_error(void);
int32_t vx_device_sync(vx_ctx* ctx);
/* raw cudaStream_t the context launches on (for CUDA-event timing by the harness) */
void* vx_ctx_stream(vx_ctx* ctx);
/* number of kernels this library has launched on this context so far */
uint64_t vx_ctx_launch_count(vx_ctx* ctx);
/* device time (ms, CUDA events on the context stream) of the phases of the most recent commit on
 * this context: out[0] staging copy, [1] iNTT, [2] coset LDE, [3] leaf hashing, [4] interior levels.
 * Plays the role of plonky2's TimingTree scopes ("IFFT", "FFT + blinding", "build Merkle tree"). */
int32_t vx_ctx_phase_ms(vx_ctx* ctx, float out[5]);

/* ---- device buffers for callers that keep prover intermediates on the GPU between calls (witness values, Z / partial
 * products, quotient coefficients: in plonky2 these are host Vecs handed from one phase of
 * prove_with_partition_witness to the next).  Every `host or device` pointer of this header may be one of these. */
int32_t vx_dev_alloc(vx_ctx* ctx, size_t bytes, uint64_t** out);
void vx_dev_free(vx_ctx* ctx, uint64_t* p);
/* dst/src: host or device, any combination; blocking */
int32_t vx_dev_copy(vx_ctx* ctx, void* dst, const void* src, size_t bytes);
/* Page-locked host memory for buffers the caller fills and hands to the library (the witness matrix that plonky2's
 * generate_partial_witness produces before prove_with_partition_witness, build.rs:69-75): a pageable source is staged by
 * the driver at ~11 GB/s, a pinned one is copied at PCIe/C2C rate.  vx_host_register pins an existing allocation in place
 * (the Rust Vec case); every pointer must be released with the matching call.  No context needed. */
int32_t vx_host_alloc(size_t bytes, void** out);
void vx_host_free(void* p);
int32_t vx_host_register(void* p, size_t bytes);
void vx_host_unregister(void* p);

/* ---- PolynomialBatch (plonky2 fri/oracle.rs) ------------------------------------------------
 * vx_commit_from_values replaces PolynomialBatch::from_values(values, rate_bits, blinding=false,
 *   cap_height, timing, fft_root_table): reached from prove_with_partition_witness
 *   (P2X/backend/circuit/build.rs:69-75, :128-134), CircuitBuilder::build
 *   (P2X/frontend/builder/mod.rs:245) and starkyx (P2X/frontend/hash/curta/stark.rs:123-126).
 *   cols: c x n column-major (polynomial j occupies cols[j*n .. (j+1)*n)), host or device.
 * vx_commit_from_coeffs replaces PolynomialBatch::from_coeffs (same, without the iNTT).
 * The batch keeps, on the device: coefficients (c x n), the LDE in leaf order (column-major,
 *   c x N with N = n << rate_bits, row j = LDE point bitrev(j)), and the Merkle digests in
 *   plonky2's interleaved layout. */
int32_t vx_commit_from_values(vx_ctx* ctx, const uint64_t* cols, uint32_t c, uint32_t log_n,
                              uint32_t rate_bits, uint32_t cap_height, vx_batch** out);
/* vx_commit_from_values that also leaves a device copy of the VALUES in values_dev_out (c x n, vx_dev_alloc): the wires
 * commit of prove_with_partition_witness, whose witness values the next phase (Z / partial products) reads again.  With
 * host `cols` the upload is the commit's own column pipeline -- no separate copy of the witness. */
int32_t vx_commit_from_values_keep(vx_ctx* ctx, const uint64_t* cols, uint32_t c, uint32_t log_n, uint32_t rate_bits,
                                   uint32_t cap_height, uint64_t* values_dev_out, vx_batch** out);
int32_t vx_commit_from_coeffs(vx_ctx* ctx, const uint64_t* coeffs, uint32_t c, uint32_t log_n,
                              uint32_t rate_bits, uint32_t cap_height, vx_batch** out);
/* The same two with the input exactly as plonky2 holds it -- `values: Vec<PolynomialValues<F>>` /
 * `polynomials: Vec<PolynomialCoeffs<F>>`, i.e. c SEPARATE heap allocations (fri/oracle.rs from_values / from_coeffs,
 * reached from P2X/backend/circuit/build.rs:69-75): cols[j] points at the n elements of column j.  No flattening on the
 * host: the columns feed the copy -> iNTT -> LDE -> leaf-sponge pipeline chunk by chunk.  The memory may be pageable
 * (a plain Vec), registered with vx_host_register, pinned, or device memory. */
int32_t vx_commit_from_values_cols(vx_ctx* ctx, const uint64_t* const* cols, uint32_t c, uint32_t log_n,
                                   uint32_t rate_bits, uint32_t cap_height, vx_batch** out);
int32_t vx_commit_from_coeffs_cols(vx_ctx* ctx, const uint64_t* const* coeffs, uint32_t c, uint32_t log_n,
                                   uint32_t rate_bits, uint32_t cap_height, vx_batch** out);
/* Multi-GPU sharding of one commit (coset partition): shard s of S (S a power of two,
 * S <= 2^cap_height) holds leaves [s*N/S, (s+1)*N/S) -- whole cap subtrees; whole cosets of the LDE while
 * S <= 2^rate_bits, and beyond that (a rate_bits = 1 STARK trace on 8 GPUs) one of the equal parts of ONE coset's leaf
 * block, which is the transform of the coefficients folded onto a smaller coset -- computed from ALL c coefficient
 * columns with no further communication.
 * Leaf indices passed to vx_batch_leaves / vx_batch_merkle_paths on a shard are LOCAL (0 .. N/S);
 * vx_batch_cap returns the shard's 2^cap_height/S cap entries. */
int32_t vx_commit_from_coeffs_shard(vx_ctx* ctx, const uint64_t* coeffs, uint32_t c, uint32_t log_n,
                                    uint32_t rate_bits, uint32_t cap_height, uint32_t shard_index,
                                    uint32_t shard_count, vx_batch** out);
/* The same partition with the exchange done by the library itself over NVLink peer memory (no NCCL on the data path):
 * every rank iNTTs its slice of the value columns, stores the coefficients into every rank's gather buffer with its
 * own kernel (release flag per rank, system scope), waits for all peers' flags, then extends and hashes its own
 * cosets / cap subtrees and pushes its cap entries to every rank the same way.  A group owns this rank's buffers:
 *   vx_shard_group_create     one per rank (rank r of `world`, same shape everywhere);
 *   vx_shard_group_ipc_handle + vx_shard_group_connect_ipc    one process per GPU: exchange the 64-byte CUDA IPC
 *                             handles out of band (world x 64 bytes, entry r = rank r's handle);
 *   vx_shard_group_connect_local   all ranks in one process (one vx_ctx per device; what a Rust prover does);
 *   vx_shard_commit_from_values    values_local = this rank's ceil(c/world) x n slice of the columns (columns
 *                             rank*ceil(c/world).., zero rows beyond c), host or device; cap_all_out (2^cap_height x 4,
 *                             host or device, may be NULL) receives the cap of the WHOLE commitment; *out is this
 *                             rank's shard, as from vx_commit_from_coeffs_shard.  ALL ranks must be inside this call
 *                             concurrently; a peer that does not show up within ~2 s makes it fail with VX_ECUDA. */
typedef struct vx_shard_group vx_shard_group;
int32_t vx_shard_group_create(vx_ctx* ctx, uint32_t rank, uint32_t world, uint32_t c, uint32_t log_n,
                              uint32_t rate_bits, uint32_t cap_height, vx_shard_group** out);
int32_t vx_shard_group_ipc_handle(vx_shard_group* g, uint8_t handle_out[64]);
int32_t vx_shard_group_connect_ipc(vx_shard_group* g, const uint8_t* handles);
int32_t vx_shard_group_connect_local(vx_shard_group* const* groups, uint32_t world);
int32_t vx_shard_commit_from_values(vx_shard_group* g, const uint64_t* values_local, uint64_t* cap_all_out,
                                    vx_batch** out);
/* this rank's gather buffer: world*ceil(c/world) x n coefficients of the last commit (device, borrowed) */
const uint64_t* vx_shard_group_coeffs_device(const vx_shard_group* g);
uint32_t vx_shard_group_cols_per_rank(const vx_shard_group* g);
void vx_shard_group_free(vx_shard_group* g);
/* Which columns a rank owns in the iNTT / exchange stage.  Small slices: rank r owns the contiguous global columns
 * [r*cpr, (r+1)*cpr) (cpr = ceil(c / world)).  Big slices (>= 256 MB received per rank, e.g. a 2^18 x 2502 STARK trace on
 * 8 GPUs): every slice is cut into equal parts and the parts of all ranks ALTERNATE in global column order, so that the
 * exchange streams behind the leaf hashing instead of every rank waiting for rank 0's slice.  vx_shard_group_column_map
 * writes, for each of this rank's cpr local columns, the global column the caller must place there (UINT32_MAX =
 * padding); vx_shard_group_set_layout overrides the choice (0 = by size, 1 = contiguous, 2 = interleaved; same value on
 * every rank, before the first commit). */
int32_t vx_shard_group_set_layout(vx_shard_group* g, uint32_t mode);
int32_t vx_shard_group_column_map(const vx_shard_group* g, uint32_t* global_col_out);
/* how long (ms) a rank waits for a peer's contribution before vx_shard_commit_from_values fails with VX_ECUDA (default
 * ~2000; 0 restores the default).  After a timeout the group is unusable -- free it and create a new one. */
int32_t vx_shard_group_set_timeout(vx_shard_group* g, uint32_t ms);
/* out[0] = first global leaf held, out[1] = leaves held, out[2] = cap entries held */
int32_t vx_batch_shard(const vx_batch* b, uint64_t out[3]);
void vx_batch_free(vx_batch* b);
/* shape: out[0]=c, out[1]=log_n, out[2]=rate_bits, out[3]=cap_height */
int32_t vx_batch_shape(const vx_batch* b, uint32_t out[4]);
/* merkle_tree.cap : 2^cap_height x 4 elements (host out) */
int32_t vx_batch_cap(vx_batch* b, uint64_t* cap_out);
/* polynomials (coefficients), c x n column-major, host or device out */
int32_t vx_batch_coeffs(vx_batch* b, uint64_t* coeffs_out);
/* merkle_tree.leaves[idx[i]] for k indices: k x c row-major (== get_lde_values(i, step) with
 * idx = bitrev(i*step)); replaces the per-query `merkle_tree.get(x)` of fri_prover_query_round */
int32_t vx_batch_leaves(vx_batch* b, const uint64_t* idx, uint32_t k, uint64_t* rows_out);
/* merkle_tree.prove(idx[i]).siblings : k x (log2 N - cap_height) x 4, bottom-up */
int32_t vx_batch_merkle_paths(vx_batch* b, const uint64_t* idx, uint32_t k, uint64_t* siblings_out);
/* full materialisation in plonky2's host layout (circuit serialisation, test_serializers at
 * P2X/backend/circuit/build.rs:282-295): leaves N x c row-major, digests 2(N - 2^cap) x 4.
 * Either pointer may be NULL. */
int32_t vx_batch_download(vx_batch* b, uint64_t* leaves_out, uint64_t* digests_out);
/* device pointers (borrowed, valid until vx_batch_free): LDE c x N column-major in leaf order,
 * coefficients c x n, digests. Used by the quotient / opening kernels and by the multi-GPU harness. */
const uint64_t* vx_batch_lde_device(const vx_batch* b);
const uint64_t* vx_batch_coeffs_device(const vx_batch* b);
const uint64_t* vx_batch_digests_device(const vx_batch* b);

/* ---- constraint evaluation (plonky2 plonk/prover.rs compute_quotient_polys +
 *      plonk/vanishing_poly.rs eval_vanishing_poly_base_batch + Gate::eval_unfiltered_base_batch;
 *      reached from P2X/backend/circuit/build.rs:69-75) ---------------------------------------
 * The circuit is described by plain data the Rust shim serialises from CommonCircuitData once per
 * circuit.  Gate constraints travel as a small register bytecode (one program for all gates, in
 * plonky2's sorted gate order); the shim obtains it by tracing Gate::eval_unfiltered_circuit, so
 * every registered gate (P2X/backend/circuit/serialization/gates.rs:85-107) is expressible.
 * Word layout: bits 0-7 opcode, 8-15 dst, 16-23 a, 24-31 b, 32-63 imm32; *_K ops read a 64-bit
 * immediate from the following word. */
enum {
    VX_OP_END = 0, VX_OP_LOADW = 1, VX_OP_LOADC = 2, VX_OP_LOADPI = 3, VX_OP_LOADK = 4,
    VX_OP_ADD = 5, VX_OP_SUB = 6, VX_OP_MUL = 7, VX_OP_ADDK = 8, VX_OP_MULK = 9, VX_OP_RSUBK = 10,
    VX_OP_SUBK = 11, VX_OP_EMIT = 12, VX_OP_BEGINGATE = 13, VX_OP_ENDGATE = 14, VX_OP_NOP = 15,
    /* Poseidon super-instructions (PoseidonGate dominates every plonky2 circuit; these run the library's native
     * Poseidon layers on interpreter registers).  SBOX7: dst = a^7.  The 12-register forms are followed by four operand
     * words: bytes 0-7 of word 1 and 0-3 of word 2 are the source registers, words 3-4 likewise the destinations
     * (all sources are read before any destination is written).
     *   MDS12K   dst = MDS * src + RC[imm]       (imm = round whose constants are added, 30 = none)
     *   DENSE12  dst = D * src + e               (MDS of full round 3 merged with the partial rounds' first matrix)
     *   PARTIAL12 src = (y, s1..s11), x0 = y + k[imm]:  dst = (25 x0 + sum v[imm][i] s_i,  s_i + w[imm][i] x0)
     * with D, e, k, v, w the tables of vx_poseidon_fast_tables. */
    VX_OP_SBOX7 = 16, VX_OP_MDS12K = 17, VX_OP_DENSE12 = 18, VX_OP_PARTIAL12 = 19,
    /* RANGE4: dst = a (a-1)(a-2)(a-3)  (2-bit limb checks of the U32 gates);  MADK: dst = a * imm64 + b (Horner steps) */
    VX_OP_RANGE4 = 20, VX_OP_MADK = 21
};
#define VX_PROGRAM_REGS 64
typedef struct vx_circuit_desc {
    uint32_t degree_bits, rate_bits;
    uint32_t num_wires, num_routed_wires;
    uint32_t num_constants;          /* columns of constants_sigmas before the sigmas (selectors + gate constants) */
    uint32_t num_selectors;
    uint32_t num_challenges;
    uint32_t num_partial_products;   /* per challenge */
    uint32_t max_degree;             /* chunk size of the permutation argument */
    uint32_t num_gate_constraints;
    const uint64_t* k_is;            /* num_routed_wires coset shifts (host) */
    const uint64_t* program;         /* gate bytecode (host) */
    uint64_t program_len;            /* in 64-bit words */
} vx_circuit_desc;

/* Z and partial-product polynomials over the trace domain (plonky2 plonk/prover.rs
 * all_wires_permutation_partial_products / wires_permutation_partial_products_and_zs).
 * wires: num_wires x n, sigmas: num_routed x n (values, host or device);
 * out: (num_challenges * (1 + num_partial_products)) x n, columns [Z_0.., pp_{0,*}, pp_{1,*}..]. */
int32_t vx_zs_partial_products(vx_ctx* ctx, const vx_circuit_desc* desc, const uint64_t* wires,
                               const uint64_t* sigmas, const uint64_t* betas, const uint64_t* gammas,
                               uint64_t* out);
/* compute_quotient_polys: evaluates the vanishing polynomial at every LDE point, divides by Z_H,
 * coset-iNTTs; quotient_coeffs_out: num_challenges x N coefficients (host or device) -- viewed as
 * (num_challenges * 2^rate_bits) x n it is the input of vx_commit_from_coeffs. */
int32_t vx_quotient(vx_ctx* ctx, const vx_circuit_desc* desc, vx_batch* constants_sigmas, vx_batch* wires,
                    vx_batch* zs_partial_products, const uint64_t pi_hash[4], const uint64_t* betas,
                    const uint64_t* gammas, const uint64_t* alphas, uint64_t* quotient_coeffs_out);
/* Circuit-load-time specialisation of the gate constraints (the counterpart of the monomorphised
 * Gate::eval_unfiltered_base_batch bodies a Rust build of plonky2 contains; a prover calls it once per circuit, next to
 * CircuitBuilder::build -- contracts/lib/succinctx/plonky2x/core/src/frontend/builder/mod.rs:245): the gate bytecode is
 * turned into straight-line CUDA and compiled for sm_100a with NVRTC (dlopen of libnvrtc.so.12, or VX_NVRTC_PATH).  From
 * then on vx_quotient runs the compiled kernel for this program on this context's device; results are bit-identical to the
 * interpreter's.  tuning: 0 = defaults; bits 0-7 minimum blocks per SM (register budget), bits 8-15 threads per block / 32,
 * bits 16-27 bytecode operations per scheduling fence.  VX_EUNSUPPORTED when NVRTC is not available -- vx_quotient then
 * keeps interpreting.  VX_QUOTIENT_JIT=0 in the environment disables the compiled path. */
int32_t vx_quotient_compile(vx_ctx* ctx, const vx_circuit_desc* desc, uint32_t tuning);
/* forget the compiled kernel of this circuit on this context's device: vx_quotient interprets again */
int32_t vx_quotient_discard(vx_ctx* ctx, const vx_circuit_desc* desc);
/* 1 if vx_quotient would run a compiled kernel for this circuit on this context, else 0 */
int32_t vx_quotient_is_compiled(vx_ctx* ctx, const vx_circuit_desc* desc);
/* The generated translation unit / its sm_100a cubin (no GPU needed: inspection and build checks).  Return the size of
 * the full result in bytes (negative: error code) and copy at most cap bytes into buf (the source NUL-terminated). */
int64_t vx_quotient_jit_source(const vx_circuit_desc* desc, char* buf, uint64_t cap);
int64_t vx_quotient_jit_cubin(const vx_circuit_desc* desc, uint32_t tuning, char* buf, uint64_t cap);
/* OpeningSet::new: every polynomial of the batch evaluated at an extension point; out: c x 2 */
int32_t vx_batch_eval_ext(vx_batch* b, const uint64_t point[2], uint64_t* out);

/* ---- FRI (plonky2 fri/oracle.rs prove_openings, fri/prover.rs fri_committed_trees /
 *      fri_proof_of_work / fri_prover_query_rounds) ---------------------------------------------
 * vx_fri_begin builds the batched opening polynomial
 *   final = sum_b alpha-shifted (sum_i alpha^i p_i - eval) / (X - point_b)
 * for the listed batches (ranges of polynomials of the given oracles), extends it over the coset and
 * keeps coefficients + values on the device.  The host challenger drives the layers:
 *   vx_fri_commit_layer -> cap (observe, draw beta) -> vx_fri_fold(beta) ... -> vx_fri_final_poly. */
typedef struct vx_fri vx_fri;
typedef struct vx_fri_range { uint32_t oracle, first, count; } vx_fri_range;
typedef struct vx_fri_batch { uint64_t point[2]; const vx_fri_range* ranges; uint32_t num_ranges; } vx_fri_batch;
int32_t vx_fri_begin(vx_ctx* ctx, vx_batch* const* oracles, uint32_t num_oracles, const vx_fri_batch* batches,
                     uint32_t num_batches, const uint64_t alpha[2], vx_fri** out);
/* Merkle tree over the current layer's bit-reversed values, `2^arity_bits` extension values per leaf */
int32_t vx_fri_commit_layer(vx_fri* f, uint32_t arity_bits, uint32_t cap_height, uint64_t* cap_out);
int32_t vx_fri_fold(vx_fri* f, const uint64_t beta[2]);
/* truncated final polynomial: out holds (len >> rate_bits) x 2 */
int32_t vx_fri_final_poly(vx_fri* f, uint64_t* out, uint32_t* len_out);
/* query openings of committed layer `layer`: leaf rows (k x 2*arity) and paths (k x depth x 4) */
int32_t vx_fri_query(vx_fri* f, uint32_t layer, const uint64_t* idx, uint32_t k, uint64_t* rows_out,
                     uint64_t* paths_out);
void vx_fri_free(vx_fri* f);
/* fri_proof_of_work: smallest w such that permute(state with state[pos] = w)[7] has >= min_zeros
 * leading zero bits (equals the reference's witness under RAYON_NUM_THREADS=1; upstream find_any
 * is nondeterministic). */
int32_t vx_pow_grind(vx_ctx* ctx, const uint64_t state[12], uint32_t pos, uint32_t min_zeros, uint64_t* witness_out);

/* ---- MerkleTree (plonky2 hash/merkle_tree.rs; API use in-tree at
 *      P2X/backend/wrapper/poseidon_bn128.rs:217-220) -----------------------------------------
 * vx_merkle_new replaces MerkleTree::<F, PoseidonHash>::new(leaves, cap_height).
 * leaves: n x w row-major, host or device.  digests_out (2(n - 2^cap) x 4, plonky2 interleaved
 * layout) and cap_out (2^cap x 4) are host buffers and may be NULL; tree_out may be NULL. */
int32_t vx_merkle_new(vx_ctx* ctx, const uint64_t* leaves, uint64_t n, uint32_t w, uint32_t cap_height,
                      uint64_t* digests_out, uint64_t* cap_out, vx_tree** tree_out);
int32_t vx_tree_prove(vx_tree* t, const uint64_t* idx, uint32_t k, uint64_t* siblings_out);
int32_t vx_tree_leaves(vx_tree* t, const uint64_t* idx, uint32_t k, uint64_t* rows_out);
int32_t vx_tree_cap(vx_tree* t, uint64_t* cap_out);
void vx_tree_free(vx_tree* t);

/* ---- the wrapper-stage hasher: PoseidonBN128Hash over GoldilocksField --------------------------------------
 * Replaces P2X/backend/wrapper/poseidon_bn128.rs:20-110 (`permution`: BN254 scalar field, t = 4, x^5, 8 + 56 rounds)
 * and the Hasher impl P2X/backend/wrapper/plonky2_config.rs:128-197 (hash_no_pad: 3 canonical Goldilocks elements =
 * 24 little-endian bytes per scalar, 3 scalars per permutation; hash_or_noop: up to 3 elements are the digest's own
 * bytes; two_to_one: permute([0, 0, l, r])[0]) for MerkleTree::<F, PoseidonBN128Hash>::new (API use and tests at
 * poseidon_bn128.rs:205-267).  A digest is one scalar = Fr::to_repr() = 32 little-endian bytes = 4 u64 words, canonical:
 * the same size as a Goldilocks-Poseidon HashOut, so `digests`, caps and paths keep their layout.
 * Known-answer test: poseidon_bn128.rs:134-181. */
#define VX_HASHER_POSEIDON 0           /* PoseidonHash (PoseidonGoldilocksConfig) */
#define VX_HASHER_POSEIDON_BN128 1     /* PoseidonBN128Hash (PoseidonBN128GoldilocksConfig) */
/* MerkleTree::<F, H>::new with H chosen by `hasher`; otherwise identical to vx_merkle_new (which is hasher 0). */
int32_t vx_merkle_new_hasher(vx_ctx* ctx, uint32_t hasher, const uint64_t* leaves, uint64_t n, uint32_t w,
                             uint32_t cap_height, uint64_t* digests_out, uint64_t* cap_out, vx_tree** tree_out);
/* PolynomialBatch::from_values / from_coeffs for a config whose Hasher is `hasher` (the wrap circuit's commitments). */
int32_t vx_commit_from_values_hasher(vx_ctx* ctx, uint32_t hasher, const uint64_t* cols, uint32_t c, uint32_t log_n,
                                     uint32_t rate_bits, uint32_t cap_height, vx_batch** out);
int32_t vx_commit_from_coeffs_hasher(vx_ctx* ctx, uint32_t hasher, const uint64_t* coeffs, uint32_t c, uint32_t log_n,
                                     uint32_t rate_bits, uint32_t cap_height, vx_batch** out);
/* `count` independent permutations: states are 4 scalars x 4 u64 words (canonical, little-endian), host in/out.
 * A word group >= the modulus is rejected with VX_EINVAL (Fr::from_repr fails in the reference). */
int32_t vx_bn128_permute(vx_ctx* ctx, const uint64_t* states_in, uint64_t count, uint64_t* states_out);
/* hash_no_pad (or_noop = 0) / hash_or_noop (or_noop = 1) of `count` inputs of `len` Goldilocks elements -> count x 4 */
int32_t vx_bn128_hash(vx_ctx* ctx, const uint64_t* inputs, uint64_t count, uint32_t len, int32_t or_noop, uint64_t* out);
/* The derived tables in the reference's layout (C_CONSTANTS[88], S_CONSTANTS[392], M_MATRIX[4][4], P_MATRIX[4][4] of
 * P2X/backend/wrapper/poseidon_bn128_constants.rs), 4 u64 words per scalar, canonical.  Host only: needs no GPU and no
 * context.  Any pointer may be NULL. */
int32_t vx_bn128_constants(uint64_t* c88, uint64_t* s392, uint64_t* m16, uint64_t* p16);

/* ---- hashing primitives (plonky2 hash/poseidon.rs, hash/hashing.rs; KAT at
 *      P2X/frontend/hash/poseidon/poseidon256.rs:163-202) --------------------------------------
 * Batched on the device: `count` independent permutations / hashes per call. Host in/out. */
int32_t vx_poseidon_permute(vx_ctx* ctx, const uint64_t* states_in, uint64_t count, uint64_t* states_out);
/* hash_n_to_hash_no_pad of `count` inputs of `len` elements each (row-major) -> count x 4 */
int32_t vx_hash_no_pad(vx_ctx* ctx, const uint64_t* inputs, uint64_t count, uint32_t len, uint64_t* out);
/* ONE permutation on the HOST (scalar C, spec round structure): the duplex step of plonky2's Challenger
 * (iop/challenger.rs), which north_star keeps on the host.  Not a fallback for the device kernels: no batch form. */
int32_t vx_challenger_permute(uint64_t state[12]);
/* the 360 round constants the device uses (for audit against the reference table) */
int32_t vx_poseidon_constants(uint64_t out[360]);
/* the derived "fast partial round" tables (same refactoring as plonky2's FAST_PARTIAL_*): dense 12x12
 * matrix D and vector e replacing the MDS layer of full round 3, then per partial round r: post-S-box
 * constant k[r], first-row entries v[r][11], first-column entries w[r][11]. Used by the PoseidonGate
 * constraint program. */
int32_t vx_poseidon_fast_tables(uint64_t dense_d[144], uint64_t dense_e[12], uint64_t k[22], uint64_t v[242],
                                uint64_t w[242]);

/* ---- field primitives (plonky2_field goldilocks_field.rs, extension/quadratic.rs), elementwise on the
 *      device over n elements (host in/out): the unit-test surface of the device arithmetic.
 *      op: 0 a*b, 1 a*b (compiler 64-bit path), 2 a^-1, 3 a*b+c, 4 a+b, 5 a-b, 6 a^7,
 *          7 extension a*b (pairs), 8 extension a^-1 (pairs; n counts pairs) */
int32_t vx_field_op(vx_ctx* ctx, uint32_t op, const uint64_t* a, const uint64_t* b, const uint64_t* c, uint64_t n,
                    uint64_t* out);

/* ---- NTT primitives (plonky2_field fft.rs) on c x n column-major batches, host or device
 *      in/out; natural order in and out ------------------------------------------------------- */
int32_t vx_ntt(vx_ctx* ctx, const uint64_t* in, uint64_t* out, uint32_t c, uint32_t log_n,
               int32_t inverse, uint64_t coset_shift /* 0 or 1 = none */);

#ifdef __cplusplus
}
#endif
#endif
