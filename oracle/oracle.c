/* See oracle.h: CPU restatement of plonky2 v0.2.0's commit path. TEST INFRASTRUCTURE ONLY. */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
#define P VXO_P
#define EPS 0xFFFFFFFFULL

/* ------------------------------------------------------------------ field */
static inline uint64_t canon(uint64_t a) { return a - ((0 - (uint64_t)(a >= P)) & P); }

static inline uint64_t add_(uint64_t a, uint64_t b) {           /* canonical in, canonical out */
    uint64_t s;
    uint64_t c = __builtin_add_overflow(a, b, &s);
    uint64_t ge = (uint64_t)(s >= P) | c;
    return s - ((0 - ge) & P);
}
static inline uint64_t sub_(uint64_t a, uint64_t b) {
    uint64_t d;
    uint64_t br = __builtin_sub_overflow(a, b, &d);
    return d + ((0 - br) & P);
}

static inline uint64_t reduce128(u128 x) {
    /* goldilocks_field.rs reduce128: lo - hi_hi + hi_lo * EPS, then canonicalise */
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & EPS;
    uint64_t t, r;                       /* branch-free: random data mispredicts badly */
    uint64_t b = __builtin_sub_overflow(lo, hh, &t);
    t -= (0 - b) & EPS;
    uint64_t m = (hl << 32) - hl;
    uint64_t c = __builtin_add_overflow(t, m, &r);
    r += (0 - c) & EPS;
    return canon(r);
}
static inline uint64_t mul_(uint64_t a, uint64_t b) { return reduce128((u128)a * b); }

uint64_t vxo_add(uint64_t a, uint64_t b) { return add_(canon(a), canon(b)); }
uint64_t vxo_sub(uint64_t a, uint64_t b) { return sub_(canon(a), canon(b)); }
uint64_t vxo_mul(uint64_t a, uint64_t b) { return mul_(a, b); }
uint64_t vxo_pow(uint64_t a, uint64_t e) {
    uint64_t r = 1, b = canon(a);
    while (e) {
        if (e & 1) r = mul_(r, b);
        b = mul_(b, b);
        e >>= 1;
    }
    return r;
}
uint64_t vxo_inv(uint64_t a) { return vxo_pow(a, P - 2); }
uint64_t vxo_root_of_unity(uint32_t log_n) {
    /* primitive_root_of_unity(k) = POWER_OF_TWO_GENERATOR^(2^(32-k)) */
    uint64_t r = 7277203076849721926ULL;
    for (uint32_t i = log_n; i < 32; i++) r = mul_(r, r);
    return r;
}

void vxo_ext_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]) {
    uint64_t a0 = canon(a[0]), a1 = canon(a[1]), b0 = canon(b[0]), b1 = canon(b[1]);
    uint64_t c0 = add_(mul_(a0, b0), mul_(7, mul_(a1, b1)));
    uint64_t c1 = add_(mul_(a0, b1), mul_(a1, b0));
    out[0] = c0;
    out[1] = c1;
}
void vxo_ext_inv(const uint64_t a[2], uint64_t out[2]) {
    /* 1/(a0 + a1 x) = (a0 - a1 x) / (a0^2 - 7 a1^2) */
    uint64_t a0 = canon(a[0]), a1 = canon(a[1]);
    uint64_t norm = sub_(mul_(a0, a0), mul_(7, mul_(a1, a1)));
    uint64_t ni = vxo_inv(norm);
    out[0] = mul_(a0, ni);
    out[1] = mul_(sub_(0, a1), ni);
}

/* ------------------------------------------------------------------ ChaCha8Rng::seed_from_u64(0) */
static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define QR(a, b, c, d)                                      \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16);           \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);           \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);            \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);

static void chacha8_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t st[16] = {0x61707865u, 0x3320646Eu, 0x79622D32u, 0x6B206574u};
    for (int i = 0; i < 8; i++) st[4 + i] = key[i];
    st[12] = (uint32_t)counter; st[13] = (uint32_t)(counter >> 32); st[14] = 0; st[15] = 0;
    uint32_t x[16];
    memcpy(x, st, sizeof x);
    for (int r = 0; r < 4; r++) {
        QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
        QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
}

static uint64_t RC[360];
static int rc_ready = 0;

static void init_constants(void) {
    if (rc_ready) return;
#pragma omp critical(vxo_rc)
    {
        if (!rc_ready) {
            /* rand_core seed_from_u64: PCG32 stream fills the 32-byte key */
            uint64_t state = 0;
            uint32_t key[8];
            for (int i = 0; i < 8; i++) {
                state = state * 6364136223846793005ULL + 11634580027462260723ULL;
                uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
                uint32_t rot = (uint32_t)(state >> 59);
                key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
            }
            uint32_t blk[16];
            int pos = 16;
            uint64_t counter = 0;
            int n = 0;
            while (n < 360) {
                uint32_t w[2];
                for (int k = 0; k < 2; k++) {
                    if (pos == 16) { chacha8_block(key, counter++, blk); pos = 0; }
                    w[k] = blk[pos++];
                }
                uint64_t v = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
                u128 m = (u128)v * P;            /* rand 0.8 sample_single: zone = P-1 */
                if ((uint64_t)m <= P - 1) RC[n++] = (uint64_t)(m >> 64);
            }
            rc_ready = 1;
        }
    }
}

void vxo_poseidon_constants(uint64_t out[360]) {
    init_constants();
    memcpy(out, RC, sizeof RC);
}

/* ------------------------------------------------------------------ Poseidon */
static const uint64_t CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};

static inline uint64_t sbox(uint64_t x) {
    uint64_t x2 = mul_(x, x), x4 = mul_(x2, x2), x3 = mul_(x, x2);
    return mul_(x3, x4);
}

static inline void mds_layer(uint64_t s[12]) {
    uint64_t o[12];
    for (int r = 0; r < 12; r++) {
        u128 acc = 0;
        for (int i = 0; i < 12; i++) acc += (u128)s[(i + r) % 12] * CIRC[i];
        if (r == 0) acc += (u128)s[0] * 8;
        o[r] = reduce128(acc);
    }
    memcpy(s, o, sizeof o);
}

void vxo_poseidon_naive(uint64_t s[12]) {
    init_constants();
    for (int i = 0; i < 12; i++) s[i] = canon(s[i]);
    for (int r = 0; r < 30; r++) {
        for (int i = 0; i < 12; i++) s[i] = add_(s[i], RC[12 * r + i]);
        if (r < 4 || r >= 26) {
            for (int i = 0; i < 12; i++) s[i] = sbox(s[i]);
        } else {
            s[0] = sbox(s[0]);
        }
        mds_layer(s);
    }
}

/* Same function, written for speed (this is the timed CPU baseline): the MDS layer is evaluated
 * on 32-bit halves with u64 accumulators (coefficients sum to 264 < 2^9), as plonky2's generic
 * mds_layer does. Checked against vxo_poseidon_naive in tests. */
static inline void mds_layer_fast(uint64_t s[12]) {
    uint64_t lo[24], hi[24];
    for (int i = 0; i < 12; i++) {
        lo[i] = lo[i + 12] = s[i] & EPS;
        hi[i] = hi[i + 12] = s[i] >> 32;
    }
#pragma GCC unroll 12
    for (int r = 0; r < 12; r++) {
        uint64_t al = 0, ah = 0;
#pragma GCC unroll 12
        for (int i = 0; i < 12; i++) {
            al += lo[i + r] * CIRC[i];
            ah += hi[i + r] * CIRC[i];
        }
        if (r == 0) { al += lo[0] * 8; ah += hi[0] * 8; }
        s[r] = reduce128((u128)al + ((u128)ah << 32));
    }
}

void vxo_poseidon(uint64_t s[12]) {
    init_constants();
    for (int i = 0; i < 12; i++) s[i] = canon(s[i]);
    for (int r = 0; r < 30; r++) {
        const uint64_t* rc = RC + 12 * r;
        for (int i = 0; i < 12; i++) s[i] = add_(s[i], rc[i]);
        if (r < 4 || r >= 26) {
            for (int i = 0; i < 12; i++) s[i] = sbox(s[i]);
        } else {
            s[0] = sbox(s[0]);
        }
        mds_layer_fast(s);
    }
}

void vxo_hash_no_pad(const uint64_t* in, size_t len, uint64_t out[4]) {
    uint64_t st[12] = {0};
    for (size_t off = 0; off < len; off += 8) {
        size_t k = len - off < 8 ? len - off : 8;
        for (size_t i = 0; i < k; i++) st[i] = canon(in[off + i]);   /* overwrite mode */
        vxo_poseidon(st);
    }
    memcpy(out, st, 4 * sizeof(uint64_t));
}

void vxo_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    uint64_t st[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
    vxo_poseidon(st);
    memcpy(out, st, 4 * sizeof(uint64_t));
}

void vxo_hash_or_noop(const uint64_t* in, size_t len, uint64_t out[4]) {
    if (len <= 4) {
        for (size_t i = 0; i < 4; i++) out[i] = i < len ? canon(in[i]) : 0;
    } else {
        vxo_hash_no_pad(in, len, out);
    }
}

/* ------------------------------------------------------------------ Merkle */
/* position of sibling pair q of layer lvl (0 = leaf digests) inside one subtree's digest block */
static inline uint64_t pair_pos(uint64_t q, uint32_t lvl) {
    return 2 * (q * (2ULL << lvl) + (1ULL << lvl) - 1);
}

void vxo_merkle_new(const uint64_t* leaves, uint64_t n, uint32_t w, uint32_t cap_height,
                    uint64_t* digests, uint64_t* cap) {
    init_constants();
    uint64_t num_caps = 1ULL << cap_height, sub = n >> cap_height;
    if (sub == 1) {
#pragma omp parallel for schedule(static)
        for (uint64_t j = 0; j < n; j++) vxo_hash_or_noop(leaves + j * w, w, cap + 4 * j);
        return;
    }
    /* layer 0: leaf digests straight into their interleaved slots */
#pragma omp parallel for schedule(static)
    for (uint64_t j = 0; j < n; j++) {
        uint64_t s = j / sub, jj = j % sub;
        uint64_t* blk = digests + 4 * s * (2 * sub - 2);
        vxo_hash_or_noop(leaves + j * w, w, blk + 4 * (pair_pos(jj >> 1, 0) + (jj & 1)));
    }
    uint32_t levels = 0;
    while ((1ULL << levels) < sub) levels++;
    for (uint32_t lvl = 0; lvl < levels; lvl++) {
        uint64_t pairs = sub >> (lvl + 1);       /* per subtree */
#pragma omp parallel for schedule(static)
        for (uint64_t t = 0; t < num_caps * pairs; t++) {
            uint64_t s = t / pairs, q = t % pairs;
            uint64_t* blk = digests + 4 * s * (2 * sub - 2);
            const uint64_t* pr = blk + 4 * pair_pos(q, lvl);
            uint64_t* dst = (lvl + 1 == levels)
                                ? cap + 4 * s
                                : blk + 4 * (pair_pos(q >> 1, lvl + 1) + (q & 1));
            vxo_two_to_one(pr, pr + 4, dst);
        }
    }
}

void vxo_merkle_prove(const uint64_t* digests, uint64_t n, uint32_t cap_height, uint64_t leaf_index,
                      uint64_t* siblings) {
    uint64_t sub = n >> cap_height;
    uint64_t s = leaf_index / sub, j = leaf_index % sub;
    const uint64_t* blk = digests + 4 * s * (2 * sub - 2);
    for (uint32_t lvl = 0; (sub >> lvl) > 1; lvl++) {
        uint64_t idx = j >> lvl;
        memcpy(siblings + 4 * lvl, blk + 4 * (pair_pos(idx >> 1, lvl) + ((idx & 1) ^ 1)),
               4 * sizeof(uint64_t));
    }
}

int vxo_merkle_verify(const uint64_t* leaf, uint32_t w, uint64_t leaf_index, const uint64_t* siblings,
                      uint32_t n_siblings, const uint64_t* cap) {
    uint64_t cur[4];
    vxo_hash_or_noop(leaf, w, cur);
    uint64_t idx = leaf_index;
    for (uint32_t i = 0; i < n_siblings; i++) {
        if (idx & 1) vxo_two_to_one(siblings + 4 * i, cur, cur);
        else vxo_two_to_one(cur, siblings + 4 * i, cur);
        idx >>= 1;
    }
    for (int i = 0; i < 4; i++)
        if (canon(cap[4 * idx + i]) != cur[i]) return 0;
    return 1;
}

/* ------------------------------------------------------------------ NTT */
static inline uint64_t bitrev64(uint64_t x, uint32_t bits) {
    uint64_t r = 0;
    for (uint32_t i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* twiddle table cache: tw[log_n][k] = w_n^k for k < n/2 (plonky2 shares one fft_root_table per
 * circuit; starkyx passes None and rebuilds it per column -- the cached variant is the faster one) */
static uint64_t* TW[33];

static const uint64_t* twiddles(uint32_t log_n) {
    if (TW[log_n]) return TW[log_n];
#pragma omp critical(vxo_tw)
    {
        if (!TW[log_n]) {
            uint64_t half = log_n ? (1ULL << (log_n - 1)) : 1;
            uint64_t* t = (uint64_t*)malloc(half * sizeof(uint64_t));
            uint64_t w = vxo_root_of_unity(log_n), acc = 1;
            for (uint64_t k = 0; k < half; k++) { t[k] = acc; acc = mul_(acc, w); }
            TW[log_n] = t;
        }
    }
    return TW[log_n];
}

static void fft_core(uint64_t* a, uint32_t log_n) {
    uint64_t n = 1ULL << log_n;
    for (uint64_t i = 0; i < n; i++) {
        a[i] = canon(a[i]);
    }
    for (uint64_t i = 0; i < n; i++) {
        uint64_t j = bitrev64(i, log_n);
        if (i < j) { uint64_t t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    if (log_n == 0) return;
    const uint64_t* tw = twiddles(log_n);
    for (uint32_t s = 1; s <= log_n; s++) {
        uint64_t m = 1ULL << s, half = m >> 1, step = n >> s;
        for (uint64_t k = 0; k < n; k += m)
            for (uint64_t j = 0; j < half; j++) {
                uint64_t u = a[k + j], v = mul_(a[k + j + half], tw[j * step]);
                a[k + j] = add_(u, v);
                a[k + j + half] = sub_(u, v);
            }
    }
}

void vxo_fft(uint64_t* buf, uint32_t log_n) { fft_core(buf, log_n); }

void vxo_ifft(uint64_t* buf, uint32_t log_n) {
    /* fft, then buf[i] <-> buf[n-i], times n^-1 */
    uint64_t n = 1ULL << log_n;
    fft_core(buf, log_n);
    uint64_t ninv = vxo_inv(n % P);
    for (uint64_t i = 1; i < n / 2; i++) { uint64_t t = buf[i]; buf[i] = buf[n - i]; buf[n - i] = t; }
    for (uint64_t i = 0; i < n; i++) buf[i] = mul_(buf[i], ninv);
}

void vxo_coset_fft(uint64_t* buf, uint32_t log_n, uint64_t shift) {
    uint64_t n = 1ULL << log_n, t = 1;
    for (uint64_t i = 0; i < n; i++) { buf[i] = mul_(buf[i], t); t = mul_(t, shift); }
    fft_core(buf, log_n);
}

void vxo_coset_ifft(uint64_t* buf, uint32_t log_n, uint64_t shift) {
    uint64_t n = 1ULL << log_n, t = 1, si = vxo_inv(shift);
    vxo_ifft(buf, log_n);
    for (uint64_t i = 0; i < n; i++) { buf[i] = mul_(buf[i], t); t = mul_(t, si); }
}

/* extension-field NTTs: the twiddles are base-field, so the two limbs transform independently */
static void split_ext(const uint64_t* buf, uint64_t n, uint64_t* a, uint64_t* b) {
    for (uint64_t i = 0; i < n; i++) { a[i] = buf[2 * i]; b[i] = buf[2 * i + 1]; }
}
static void join_ext(uint64_t* buf, uint64_t n, const uint64_t* a, const uint64_t* b) {
    for (uint64_t i = 0; i < n; i++) { buf[2 * i] = a[i]; buf[2 * i + 1] = b[i]; }
}
void vxo_fft_ext(uint64_t* buf, uint32_t log_n) {
    uint64_t n = 1ULL << log_n;
    uint64_t* a = (uint64_t*)malloc(2 * n * sizeof(uint64_t));
    split_ext(buf, n, a, a + n);
    fft_core(a, log_n); fft_core(a + n, log_n);
    join_ext(buf, n, a, a + n);
    free(a);
}
void vxo_coset_fft_ext(uint64_t* buf, uint32_t log_n, const uint64_t shift) {
    uint64_t n = 1ULL << log_n;
    uint64_t* a = (uint64_t*)malloc(2 * n * sizeof(uint64_t));
    split_ext(buf, n, a, a + n);
    vxo_coset_fft(a, log_n, shift); vxo_coset_fft(a + n, log_n, shift);
    join_ext(buf, n, a, a + n);
    free(a);
}
void vxo_ifft_ext(uint64_t* buf, uint32_t log_n) {
    uint64_t n = 1ULL << log_n;
    uint64_t* a = (uint64_t*)malloc(2 * n * sizeof(uint64_t));
    split_ext(buf, n, a, a + n);
    vxo_ifft(a, log_n); vxo_ifft(a + n, log_n);
    join_ext(buf, n, a, a + n);
    free(a);
}

/* ------------------------------------------------------------------ commit */
void vxo_commit_from_coeffs(const uint64_t* coeffs, uint32_t c, uint32_t log_n, uint32_t rate_bits,
                            uint32_t cap_height, uint64_t* leaves, uint64_t* digests, uint64_t* cap) {
    uint64_t n = 1ULL << log_n, N = n << rate_bits;
    uint32_t log_N = log_n + rate_bits;
    uint64_t* lde = (uint64_t*)malloc((size_t)c * N * sizeof(uint64_t));   /* column-major, as Vec<Vec<F>> */
    int own_leaves = leaves == NULL;
    if (own_leaves) leaves = (uint64_t*)malloc((size_t)c * N * sizeof(uint64_t));
    twiddles(log_N);
    /* "FFT + blinding" (blinding = false): per-column lde + coset_fft(g), one column per task */
#pragma omp parallel for schedule(dynamic, 1)
    for (uint32_t j = 0; j < c; j++) {
        uint64_t* col = lde + (size_t)j * N;
        memcpy(col, coeffs + (size_t)j * n, n * sizeof(uint64_t));
        memset(col + n, 0, (N - n) * sizeof(uint64_t));
        vxo_coset_fft(col, log_N, 14293326489335486720ULL);
    }
    /* "transpose LDEs" + reverse_index_bits_in_place */
#pragma omp parallel for schedule(static)
    for (uint64_t r = 0; r < N; r++) {
        uint64_t i = bitrev64(r, log_N);
        uint64_t* row = leaves + r * c;
        for (uint32_t j = 0; j < c; j++) row[j] = lde[(size_t)j * N + i];
    }
    free(lde);
    /* "build Merkle tree" */
    int own_dig = digests == NULL;
    if (own_dig) digests = (uint64_t*)malloc((size_t)8 * N * sizeof(uint64_t) + 64);
    vxo_merkle_new(leaves, N, c, cap_height, digests, cap);
    if (own_dig) free(digests);
    if (own_leaves) free(leaves);
}

void vxo_commit_from_values(const uint64_t* cols, uint32_t c, uint32_t log_n, uint32_t rate_bits,
                            uint32_t cap_height, uint64_t* coeffs, uint64_t* leaves,
                            uint64_t* digests, uint64_t* cap) {
    uint64_t n = 1ULL << log_n;
    int own = coeffs == NULL;
    if (own) coeffs = (uint64_t*)malloc((size_t)c * n * sizeof(uint64_t));
    twiddles(log_n);
    /* "IFFT": one column per task */
#pragma omp parallel for schedule(dynamic, 1)
    for (uint32_t j = 0; j < c; j++) {
        uint64_t* col = coeffs + (size_t)j * n;
        memcpy(col, cols + (size_t)j * n, n * sizeof(uint64_t));
        vxo_ifft(col, log_n);
    }
    vxo_commit_from_coeffs(coeffs, c, log_n, rate_bits, cap_height, leaves, digests, cap);
    if (own) free(coeffs);
}

int vxo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void vxo_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
