/* CPU oracle for the plonky2 v0.2.0 proving hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libvectorx_b200.so) never links, loads or calls it.
 *
 * The arithmetic of the path lives in third-party crates that are NOT vendored under
 * /root/reference (plonky2 + plonky2_field v0.2.0 @ 7445ec911b0c, Cargo.lock:4847-4850,4871-4873);
 * every function restates the published algorithm and cites the reference
 * call site that reaches it.  Pinned by the reference's only golden vector at this boundary
 * (P2X/frontend/hash/poseidon/poseidon256.rs:163-202) -- everything downstream of Poseidon
 * (LDE values, caps, quotients, FRI) is "parity unpinned" by reference data and is instead held
 * by algebraic self-checks and the independent big-int restatement in oracle/pyref.py.
 *
 * P2X = contracts/lib/succinctx/plonky2x/core/src
 */
#ifndef VX_ORACLE_H
#define VX_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VXO_P 0xFFFFFFFF00000001ULL

/* ---- field (plonky2_field goldilocks_field.rs; type bound at P2X/backend/circuit/config.rs:37) */
uint64_t vxo_add(uint64_t a, uint64_t b);
uint64_t vxo_sub(uint64_t a, uint64_t b);
uint64_t vxo_mul(uint64_t a, uint64_t b);
uint64_t vxo_pow(uint64_t a, uint64_t e);
uint64_t vxo_inv(uint64_t a);
uint64_t vxo_root_of_unity(uint32_t log_n);
/* quadratic extension x^2 = 7 (config.rs:41), a = a[0] + a[1] x */
void vxo_ext_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]);
void vxo_ext_inv(const uint64_t a[2], uint64_t out[2]);

/* ---- Poseidon (plonky2 hash/poseidon.rs, poseidon_goldilocks.rs; KAT poseidon256.rs:163-202) */
void vxo_poseidon_constants(uint64_t out[360]);
void vxo_poseidon(uint64_t state[12]);
void vxo_poseidon_naive(uint64_t state[12]);     /* literal spec form, used to check vxo_poseidon */
void vxo_hash_no_pad(const uint64_t* in, size_t len, uint64_t out[4]);   /* P2X/utils/poseidon/mod.rs:31-36 */
void vxo_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);
void vxo_hash_or_noop(const uint64_t* in, size_t len, uint64_t out[4]);

/* ---- MerkleTree::new (plonky2 hash/merkle_tree.rs; API exercised at
 *      P2X/backend/wrapper/poseidon_bn128.rs:217-220).  leaves: n x w row-major.
 *      digests: 2*(n - 2^cap) x 4 in plonky2's interleaved layout; cap: 2^cap x 4. */
void vxo_merkle_new(const uint64_t* leaves, uint64_t n, uint32_t w, uint32_t cap_height,
                    uint64_t* digests, uint64_t* cap);
/* siblings: (log2 n - cap_height) x 4, bottom-up */
void vxo_merkle_prove(const uint64_t* digests, uint64_t n, uint32_t cap_height, uint64_t leaf_index,
                      uint64_t* siblings);
int vxo_merkle_verify(const uint64_t* leaf, uint32_t w, uint64_t leaf_index, const uint64_t* siblings,
                      uint32_t n_siblings, const uint64_t* cap);

/* ---- NTT (plonky2_field fft.rs), in place, natural order in and out */
void vxo_fft(uint64_t* buf, uint32_t log_n);
void vxo_ifft(uint64_t* buf, uint32_t log_n);
void vxo_coset_fft(uint64_t* buf, uint32_t log_n, uint64_t shift);
void vxo_coset_ifft(uint64_t* buf, uint32_t log_n, uint64_t shift);
/* extension-field variants: buf holds n pairs (limb0, limb1) */
void vxo_fft_ext(uint64_t* buf, uint32_t log_n);
void vxo_coset_fft_ext(uint64_t* buf, uint32_t log_n, const uint64_t shift);
void vxo_ifft_ext(uint64_t* buf, uint32_t log_n);

/* ---- PolynomialBatch::{from_values,from_coeffs} (plonky2 fri/oracle.rs; reached from
 *      P2X/backend/circuit/build.rs:69-75 and P2X/frontend/hash/curta/stark.rs:123-126).
 *      cols / coeffs: c x n column-major.  leaves: N x c row-major, N = n << rate_bits, row j =
 *      LDE point bitrev(j).  digests/cap as vxo_merkle_new.  Any out pointer may be NULL except cap. */
void vxo_commit_from_values(const uint64_t* cols, uint32_t c, uint32_t log_n, uint32_t rate_bits,
                            uint32_t cap_height, uint64_t* coeffs, uint64_t* leaves,
                            uint64_t* digests, uint64_t* cap);
void vxo_commit_from_coeffs(const uint64_t* coeffs, uint32_t c, uint32_t log_n, uint32_t rate_bits,
                            uint32_t cap_height, uint64_t* leaves, uint64_t* digests, uint64_t* cap);

int vxo_num_threads(void);
void vxo_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
