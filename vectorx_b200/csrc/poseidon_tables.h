// Host-side derivation of every Poseidon constant table the kernels use.
//
// Naive tables: round constants RC[30][12] (ChaCha8Rng::seed_from_u64(0), see merkle.cu) and the
// MDS matrix M[r][c] = CIRC[(c - r) mod 12] + DIAG[r]*[r==c].
//
// "Fast partial rounds" (the same refactoring plonky2 v0.2.0 ships as FAST_PARTIAL_* in
// hash/poseidon_goldilocks.rs; derived here algebraically, not copied):
//   the 22 partial rounds  x <- M * S0(x + c_r)   (S0 = x^7 on lane 0 only) are rewritten as
//     x <- INIT * (x + first)                                   once, then per round r
//     x0 <- x0^7 + k[r];  d = M00*x0 + sum_i v[r][i]*x_i;  x_i += w[r][i]*x0 (i>=1);  x0 <- d
//   by pushing lane>=1 constants backwards through M^-1 and factoring A = M''*M' with
//   M' = diag(1, A[1:,1:]) (commutes with lane-0 operations) and M'' sparse (first row/column).
// The dense INIT is merged with the preceding full round's MDS layer: D = INIT*M, e = INIT*first.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

struct PoseidonTables {
    unsigned long long rc[31 * 12];        // RC[30][12] + one zero row
    unsigned int rc22[31 * 12 * 3];        // the same constants as 22/22/20-bit limbs (variant S MDS layer)
    unsigned long long dense_d[12 * 12];   // D = INIT * M   (replaces the MDS layer of full round 3)
    unsigned long long dense_e[12];        // e = INIT * first
    unsigned long long pk[22];             // post-S-box lane-0 constants (last is 0)
    unsigned long long pv[22 * 11];        // first-row entries  (d = 25*x0 + sum v_i x_i)
    unsigned long long pw[22 * 11];        // first-column entries (x_i += w_i x0)
    // grouped partial rounds: pc[r][q] = sum_i v[r][i] w[q][i] for q < r, the weight of round q's S-box output in round
    // r's first-row dot product when lanes 1..11 are only brought up to date once per group of rounds
    unsigned long long pc[22 * 22];
};

#ifndef __CUDACC_RTC__
// fills t from the 360 round constants; returns false if a matrix was singular (never happens)
bool poseidon_derive_tables(const unsigned long long rc360[360], PoseidonTables* t);
// hybrid form: the first `naive` (0..21) partial rounds stay in the spec form; pk/pv/pw hold 22 - naive rounds
bool poseidon_derive_tables_hybrid(const unsigned long long rc360[360], int naive, PoseidonTables* t);
#endif
