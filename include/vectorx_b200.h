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

typedef struct vx_ctx vx_ctx;       /* per-device context: streams, twiddles, constants, pools */
typedef struct vx_batch vx_batch;   /* device-resident PolynomialBatch (coeffs + LDE + Merkle tree) */
typedef struct vx_tree vx_tree;     /* device-resident MerkleTree over row-major leaves */

/* ---- context ------------------------------------------------------------------------------ */
int32_t vx_ctx_create(int32_t device, vx_ctx** out);
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
int32_t vx_commit_from_coeffs(vx_ctx* ctx, const uint64_t* coeffs, uint32_t c, uint32_t log_n,
                              uint32_t rate_bits, uint32_t cap_height, vx_batch** out);
/* Multi-GPU sharding of one commit (coset partition): shard s of S (S a power of two,
 * S <= 2^rate_bits, S <= 2^cap_height) holds leaves [s*N/S, (s+1)*N/S) -- whole cosets of the LDE
 * and whole cap subtrees -- computed from ALL c coefficient columns with no further communication.
 * Leaf indices passed to vx_batch_leaves / vx_batch_merkle_paths on a shard are LOCAL (0 .. N/S);
 * vx_batch_cap returns the shard's 2^cap_height/S cap entries. */
int32_t vx_commit_from_coeffs_shard(vx_ctx* ctx, const uint64_t* coeffs, uint32_t c, uint32_t log_n,
                                    uint32_t rate_bits, uint32_t cap_height, uint32_t shard_index,
                                    uint32_t shard_count, vx_batch** out);
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

/* ---- hashing primitives (plonky2 hash/poseidon.rs, hash/hashing.rs; KAT at
 *      P2X/frontend/hash/poseidon/poseidon256.rs:163-202) --------------------------------------
 * Batched on the device: `count` independent permutations / hashes per call. Host in/out. */
int32_t vx_poseidon_permute(vx_ctx* ctx, const uint64_t* states_in, uint64_t count, uint64_t* states_out);
/* hash_n_to_hash_no_pad of `count` inputs of `len` elements each (row-major) -> count x 4 */
int32_t vx_hash_no_pad(vx_ctx* ctx, const uint64_t* inputs, uint64_t count, uint32_t len, uint64_t* out);
/* the 360 round constants the device uses (for audit against the reference table) */
int32_t vx_poseidon_constants(uint64_t out[360]);

/* ---- NTT primitives (plonky2_field fft.rs) on c x n column-major batches, host or device
 *      in/out; natural order in and out ------------------------------------------------------- */
int32_t vx_ntt(vx_ctx* ctx, const uint64_t* in, uint64_t* out, uint32_t c, uint32_t log_n,
               int32_t inverse, uint64_t coset_shift /* 0 or 1 = none */);

#ifdef __cplusplus
}
#endif
#endif
