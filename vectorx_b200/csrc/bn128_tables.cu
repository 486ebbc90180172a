// PoseidonBN128 tables, derived on the host (see bn128_tables.h).  Host-only code: Montgomery arithmetic on 4 x 64-bit limbs
// with unsigned __int128, the Poseidon reference generator's Grain LFSR, Gauss-Jordan inversion.
#include "bn128_tables.h"

#include <cstring>
#include <vector>

namespace {
typedef unsigned __int128 u128;
typedef Bn128Fr Fr;

// r = 21888242871839275222246405745257275088548364400416034343698204186575808495617
// (PrimeFieldModulus in contracts/lib/succinctx/plonky2x/core/src/backend/wrapper/utils.rs:4)
const Fr N = {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};

bool geq(const Fr& a, const Fr& b) {
    for (int i = 3; i >= 0; i--)
        if (a.l[i] != b.l[i]) return a.l[i] > b.l[i];
    return true;
}
Fr sub_raw(const Fr& a, const Fr& b) {
    Fr r; u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.l[i] - b.l[i] - borrow;
        r.l[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    return r;
}
Fr add(const Fr& a, const Fr& b) {           // a, b < N < 2^254: no overflow out of 256 bits
    Fr r; u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    return geq(r, N) ? sub_raw(r, N) : r;
}
Fr sub(const Fr& a, const Fr& b) {           // a - b mod N
    if (geq(a, b)) return sub_raw(a, b);
    Fr t; u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + N.l[i]; t.l[i] = (uint64_t)c; c >>= 64; }
    return sub_raw(t, b);
}

uint64_t n0inv() {                           // -N^-1 mod 2^64 (Newton)
    uint64_t x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - N.l[0] * x;
    return ~x + 1;
}
const uint64_t N0INV = n0inv();

Fr mont_mul(const Fr& a, const Fr& b) {      // a b 2^-256 mod N (CIOS)
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * N0INV;
        c = (u128)m * N.l[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * N.l[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fr r = {{t[0], t[1], t[2], t[3]}};
    return (t[4] || geq(r, N)) ? sub_raw(r, N) : r;
}
Fr r2() {                                    // 2^512 mod N by doubling
    Fr x = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; i++) x = add(x, x);
    return x;
}
const Fr R2 = r2();
const Fr ONE_RAW = {{1, 0, 0, 0}};
Fr to_mont(const Fr& a) { return mont_mul(a, R2); }
Fr from_mont(const Fr& a) { return mont_mul(a, ONE_RAW); }
const Fr ZERO = {{0, 0, 0, 0}};
const Fr ONE = to_mont(ONE_RAW);
bool is_zero(const Fr& a) { return !(a.l[0] | a.l[1] | a.l[2] | a.l[3]); }

Fr inv(const Fr& a) {                        // a^(N-2), Montgomery form in and out
    Fr e = sub_raw(N, {{2, 0, 0, 0}});
    Fr r = ONE, b = a;
    for (int i = 0; i < 256; i++) {
        if ((e.l[i / 64] >> (i % 64)) & 1) r = mont_mul(r, b);
        b = mont_mul(b, b);
    }
    return r;
}

typedef std::vector<std::vector<Fr>> Mat;
typedef std::vector<Fr> Vec;
Vec matvec(const Mat& A, const Vec& x) {
    Vec y(A.size(), ZERO);
    for (size_t i = 0; i < A.size(); i++)
        for (size_t j = 0; j < x.size(); j++) y[i] = add(y[i], mont_mul(A[i][j], x[j]));
    return y;
}
Mat matmul(const Mat& A, const Mat& B) {
    Mat C(A.size(), Vec(B[0].size(), ZERO));
    for (size_t i = 0; i < A.size(); i++)
        for (size_t j = 0; j < B[0].size(); j++)
            for (size_t k = 0; k < B.size(); k++) C[i][j] = add(C[i][j], mont_mul(A[i][k], B[k][j]));
    return C;
}
bool invert(const Mat& A, Mat& out) {
    size_t n = A.size();
    Mat M(n, Vec(2 * n, ZERO));
    for (size_t i = 0; i < n; i++) {
        for (size_t j = 0; j < n; j++) M[i][j] = A[i][j];
        M[i][n + i] = ONE;
    }
    for (size_t c = 0; c < n; c++) {
        size_t p = c;
        while (p < n && is_zero(M[p][c])) p++;
        if (p == n) return false;
        std::swap(M[c], M[p]);
        Fr iv = inv(M[c][c]);
        for (auto& x : M[c]) x = mont_mul(x, iv);
        for (size_t r = 0; r < n; r++) {
            if (r == c || is_zero(M[r][c])) continue;
            Fr f = M[r][c];
            for (size_t j = 0; j < 2 * n; j++) M[r][j] = sub(M[r][j], mont_mul(f, M[c][j]));
        }
    }
    out.assign(n, Vec(n));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) out[i][j] = M[i][n + j];
    return true;
}

// The Poseidon reference generator's Grain LFSR (80 bits, self-shrinking output):
// field = 1 (prime), sbox = 0 (x^alpha), n = 254, t = 4, R_F = 8, R_P = 56, thirty ones; 160 warm-up steps.
struct Grain {
    uint8_t st[80];
    int head = 0;
    Grain() {
        int k = 0;
        auto push = [&](unsigned v, int w) { for (int i = w - 1; i >= 0; i--) st[k++] = (v >> i) & 1; };
        push(1, 2); push(0, 4); push(254, 12); push(4, 12); push(8, 10); push(56, 10); push(0x3fffffffu, 30);
        for (int i = 0; i < 160; i++) step();
    }
    int at(int i) const { return st[(head + i) % 80]; }
    int step() {
        int nb = at(62) ^ at(51) ^ at(38) ^ at(23) ^ at(13) ^ at(0);
        st[head] = (uint8_t)nb;
        head = (head + 1) % 80;
        return nb;
    }
    int bit() {
        for (;;) { int a = step(), b = step(); if (a) return b; }
    }
    Fr sample() {                            // 254 bits, most significant first
        Fr v = ZERO;
        for (int i = 253; i >= 0; i--) v.l[i / 64] |= (uint64_t)bit() << (i % 64);
        return v;
    }
    Fr field() {                             // rejection sampling (round constants)
        for (;;) { Fr v = sample(); if (!geq(v, N)) return v; }
    }
};
}  // namespace

Bn128Fr bn128_to_mont(const Bn128Fr& a) { return to_mont(a); }
Bn128Fr bn128_from_mont(const Bn128Fr& a) { return from_mont(a); }

bool bn128_derive_tables(Bn128Tables* out) {
    const int T = 4, RF = 8, RP = 56, HALF = RF / 2, LAST_PARTIAL = HALF + RP - 1;
    Grain g;
    std::vector<Vec> c(RF + RP, Vec(T));
    for (auto& row : c)
        for (auto& x : row) x = to_mont(g.field());
    Vec xy(2 * T);
    for (auto& v : xy) {                     // no rejection here: the generator reduces mod r
        Fr s = g.sample();
        v = to_mont(geq(s, N) ? sub_raw(s, N) : s);
    }
    Mat mds(T, Vec(T));
    for (int i = 0; i < T; i++)
        for (int j = 0; j < T; j++) mds[i][j] = inv(add(xy[i], xy[T + j]));
    Mat minv;
    if (!invert(mds, minv)) return false;
    // constants: c_r added after a partial round's matrix == M^-1 c_r added before it; lane 0 stays behind that round's
    // S-box, lanes 1..3 commute with the partial S-box and join the round's own constants
    Vec post(RF + RP, ZERO);
    for (int r = LAST_PARTIAL + 1; r > HALF; r--) {
        Vec cp = matvec(minv, c[r]);
        post[r - 1] = cp[0];
        for (int k = 1; k < T; k++) c[r - 1][k] = add(c[r - 1][k], cp[k]);
    }
    std::vector<Fr> C;
    for (auto& x : c[0]) C.push_back(x);
    for (int r = 1; r <= HALF; r++)
        for (auto& x : matvec(minv, c[r])) C.push_back(x);
    for (int r = HALF; r <= LAST_PARTIAL; r++) C.push_back(post[r]);
    for (int r = LAST_PARTIAL + 2; r < RF + RP; r++)
        for (auto& x : matvec(minv, c[r])) C.push_back(x);
    if (C.size() != 88) return false;
    // matrices: M_mul = sparse x blockdiag(1, M_hat), last partial round first
    std::vector<Fr> S(7 * RP);
    Mat mmul = mds;
    for (int i = RP - 1; i >= 0; i--) {
        Mat mhat(T - 1, Vec(T - 1)), mhi;
        for (int a = 1; a < T; a++)
            for (int b = 1; b < T; b++) mhat[a - 1][b - 1] = mmul[a][b];
        if (!invert(mhat, mhi)) return false;
        S[7 * i] = mmul[0][0];
        for (int j = 0; j < T - 1; j++) {
            Fr v = ZERO;
            for (int k = 0; k < T - 1; k++) v = add(v, mont_mul(mmul[0][1 + k], mhi[k][j]));
            S[7 * i + 1 + j] = v;
            S[7 * i + T + j] = mmul[1 + j][0];
        }
        Mat mp(T, Vec(T, ZERO));
        mp[0][0] = ONE;
        for (int a = 1; a < T; a++)
            for (int b = 1; b < T; b++) mp[a][b] = mhat[a - 1][b - 1];
        mmul = matmul(mp, mds);
    }
    for (int i = 0; i < 88; i++) out->C[i] = from_mont(C[i]);
    for (int i = 0; i < 7 * RP; i++) out->S[i] = from_mont(S[i]);
    for (int j = 0; j < T; j++)
        for (int i = 0; i < T; i++) {        // stored transposed: mix() reads constant_matrix[j][i]
            out->M[4 * j + i] = from_mont(mds[i][j]);
            out->P[4 * j + i] = from_mont(mmul[i][j]);
        }
    return true;
}
