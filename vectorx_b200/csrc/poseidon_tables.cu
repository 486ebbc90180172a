#include "poseidon_tables.h"

#include <cstring>
#include <vector>

namespace {
typedef unsigned __int128 u128;
const uint64_t P = 0xFFFFFFFF00000001ULL;
inline uint64_t mulm(uint64_t a, uint64_t b) { return (uint64_t)((u128)a * b % P); }
inline uint64_t addm(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a + b) % P); }
inline uint64_t subm(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a + P - b % P) % P); }
uint64_t powm(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    while (e) { if (e & 1) r = mulm(r, a); a = mulm(a, a); e >>= 1; }
    return r;
}
typedef std::vector<std::vector<uint64_t>> Mat;

Mat matmul(const Mat& A, const Mat& B) {
    size_t n = A.size(), k = B.size(), m = B[0].size();
    Mat C(n, std::vector<uint64_t>(m, 0));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < m; j++) {
            uint64_t acc = 0;
            for (size_t t = 0; t < k; t++) acc = addm(acc, mulm(A[i][t], B[t][j]));
            C[i][j] = acc;
        }
    return C;
}
std::vector<uint64_t> matvec(const Mat& A, const std::vector<uint64_t>& x) {
    std::vector<uint64_t> y(A.size(), 0);
    for (size_t i = 0; i < A.size(); i++)
        for (size_t j = 0; j < x.size(); j++) y[i] = addm(y[i], mulm(A[i][j], x[j]));
    return y;
}
bool invert(const Mat& A, Mat& out) {       // Gauss-Jordan mod p
    size_t n = A.size();
    Mat M(n, std::vector<uint64_t>(2 * n, 0));
    for (size_t i = 0; i < n; i++) {
        for (size_t j = 0; j < n; j++) M[i][j] = A[i][j] % P;
        M[i][n + i] = 1;
    }
    for (size_t c = 0; c < n; c++) {
        size_t p = c;
        while (p < n && M[p][c] == 0) p++;
        if (p == n) return false;
        std::swap(M[c], M[p]);
        uint64_t iv = powm(M[c][c], P - 2);
        for (auto& x : M[c]) x = mulm(x, iv);
        for (size_t r = 0; r < n; r++) {
            if (r == c || M[r][c] == 0) continue;
            uint64_t f = M[r][c];
            for (size_t j = 0; j < 2 * n; j++) M[r][j] = subm(M[r][j], mulm(f, M[c][j]));
        }
    }
    out.assign(n, std::vector<uint64_t>(n));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) out[i][j] = M[i][n + j];
    return true;
}
}  // namespace

bool poseidon_derive_tables(const unsigned long long rc360[360], PoseidonTables* t) {
    return poseidon_derive_tables_hybrid(rc360, 0, t);
}

// The same derivation for the LAST 22 - naive partial rounds only (rounds 4 + naive .. 25): the first `naive` partial
// rounds then keep the spec form (x0^7, MDS, constants) and the dense layer replaces the MDS of round 3 + naive.
bool poseidon_derive_tables_hybrid(const unsigned long long rc360[360], int naive, PoseidonTables* t) {
    static const uint64_t CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    if (naive < 0 || naive > 21) return false;
    const int R = 22 - naive, FIRST_PARTIAL = 4 + naive;
    memset(t, 0, sizeof *t);
    for (int i = 0; i < 360; i++) {
        t->rc[i] = rc360[i];
        t->rc22[3 * i] = (unsigned)(rc360[i] & 0x3fffff);
        t->rc22[3 * i + 1] = (unsigned)((rc360[i] >> 22) & 0x3fffff);
        t->rc22[3 * i + 2] = (unsigned)(rc360[i] >> 44);
    }
    Mat M(12, std::vector<uint64_t>(12));
    for (int r = 0; r < 12; r++)
        for (int c = 0; c < 12; c++) M[r][c] = CIRC[(c - r + 12) % 12] + ((r == 0 && c == 0) ? 8 : 0);
    Mat Minv;
    if (!invert(M, Minv)) return false;
    // constants, backwards: a = vector that must be added before the S-box of round r
    std::vector<uint64_t> a(12); for (int i = 0; i < 12; i++) a[i] = rc360[12 * (FIRST_PARTIAL + R - 1) + i];
    t->pk[R - 1] = 0;
    for (int r = R - 1; r >= 1; r--) {
        std::vector<uint64_t> b = matvec(Minv, a);        // same thing added after the S-box of round r-1
        t->pk[r - 1] = b[0];
        const unsigned long long* c = rc360 + 12 * (FIRST_PARTIAL + r - 1);
        a[0] = c[0];
        for (int i = 1; i < 12; i++) a[i] = addm(c[i], b[i]);
    }
    std::vector<uint64_t> first = a;
    // matrices, backwards: A = M' * M = M'' * M'_new
    Mat Mp(12, std::vector<uint64_t>(12, 0));
    for (int i = 0; i < 12; i++) Mp[i][i] = 1;
    for (int r = R - 1; r >= 0; r--) {
        Mat A = matmul(Mp, M);
        Mat Ah(11, std::vector<uint64_t>(11)), Ahi;
        for (int i = 0; i < 11; i++)
            for (int j = 0; j < 11; j++) Ah[i][j] = A[i + 1][j + 1];
        if (!invert(Ah, Ahi)) return false;
        for (int i = 0; i < 11; i++) t->pw[r * 11 + i] = A[i + 1][0];
        for (int j = 0; j < 11; j++) {
            uint64_t acc = 0;
            for (int k = 0; k < 11; k++) acc = addm(acc, mulm(A[0][k + 1], Ahi[k][j]));
            t->pv[r * 11 + j] = acc;
        }
        for (auto& row : Mp) std::fill(row.begin(), row.end(), 0);
        Mp[0][0] = 1;
        for (int i = 0; i < 11; i++)
            for (int j = 0; j < 11; j++) Mp[i + 1][j + 1] = Ah[i][j];
    }
    for (int r = 0; r < R; r++)
        for (int q = 0; q < r; q++) {
            uint64_t acc = 0;
            for (int i = 0; i < 11; i++) acc = addm(acc, mulm(t->pv[r * 11 + i], t->pw[q * 11 + i]));
            t->pc[r * 22 + q] = acc;
        }
    Mat D = matmul(Mp, M);
    std::vector<uint64_t> e = matvec(Mp, first);
    for (int i = 0; i < 12; i++) {
        t->dense_e[i] = e[i];
        for (int j = 0; j < 12; j++) t->dense_d[i * 12 + j] = D[i][j];
    }
    return true;
}
