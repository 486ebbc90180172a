// Host check of vectorx_b200/csrc/ntt_l3.cuh (compiled by tests/test_ntt_l3_host.py with g++, no CUDA): the lazily reduced
// 3-limb arithmetic -- folds through phi = 2^32, products by 2^s, the counter growth bound -- and the 2^R-point DIF network,
// forward and inverse, against a naive DFT over Goldilocks with w = 2^(192 / 2^R).
#include "ntt_l3.cuh"

#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
typedef unsigned __int128 u128;
static const u64 P = 0xFFFFFFFF00000001ULL;
static u64 mulm(u64 a, u64 b) { return (u64)((u128)a * b % P); }
static u64 powm(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = mulm(r, a); a = mulm(a, a); e >>= 1; } return r; }
static int brev(int q, int bits) { int r = 0; for (int i = 0; i < bits; i++) r |= ((q >> i) & 1) << (bits - 1 - i); return r; }
static u64 rnd() {
    switch (rand() % 8) {
        case 0: return 0;
        case 1: return P - 1;
        case 2: return 0xFFFFFFFFFFFFFFFFULL;       // non-canonical representative
        case 3: return 0xFFFFFFFFULL;
        case 4: return P;                           // non-canonical zero
        default: return ((u64)rand() << 40) ^ ((u64)rand() << 20) ^ (u64)rand();
    }
}
template <int R, bool INV>
static int run(int trials) {
    const int n = 1 << R;
    u64 w = powm(2, 192 / n);
    if (INV) w = powm(w, P - 2);
    for (int t = 0; t < trials; t++) {
        u64 in[16];
        L3 x[16];
        for (int k = 0; k < n; k++) { in[k] = rnd(); x[k] = l3_from(in[k]); }
        l3_dft<R, INV>(x);
        for (int q = 0; q < n; q++) {
            const int j = brev(q, R);
            u64 acc = 0;
            for (int k = 0; k < n; k++) acc = (u64)(((u128)acc + mulm(in[k] % P, powm(w, (u64)j * k))) % P);
            if (abs(x[q].c) > 23) { printf("R=%d: counter %d out of the documented bound\n", R, x[q].c); return 1; }
            if (l3_reduce(x[q]) != acc) { printf("R=%d inv=%d: output %d differs\n", R, (int)INV, q); return 1; }
        }
    }
    return 0;
}
template <int SH>
static int shifts(int trials) {
    for (int t = 0; t < trials; t++) {
        L3 x = l3_from(rnd());
        // the documented precondition |c| < 2^(30 - s): 11 is the largest counter a shifting stage sees (stage 2, s = 16);
        // the s = 28 shifts of stage 0 see |c| <= 1
        const int lim = (30 - (SH & 31)) >= 5 ? 11 : (1 << (30 - (SH & 31))) - 1;
        x.c = rand() % (2 * lim + 1) - lim;
        const __int128 v = l3_val(x);
        const L3 y = l3_mul2exp<SH>(x);
        __int128 want = ((v % (__int128)P) + (__int128)P) % (__int128)P;
        want = (__int128)mulm((u64)want, powm(2, SH));
        if (l3_reduce(y) != (u64)want || abs(y.c) > 2) { printf("2^%d: wrong product or counter %d\n", SH, y.c); return 1; }
    }
    return 0;
}
int main() {
    int bad = 0;
    bad |= run<1, false>(500) | run<2, false>(500) | run<3, false>(500) | run<4, false>(3000);
    bad |= run<1, true>(500) | run<2, true>(500) | run<3, true>(500) | run<4, true>(3000);
    bad |= shifts<4>(2000) | shifts<12>(2000) | shifts<16>(2000) | shifts<24>(2000) | shifts<28>(2000) | shifts<32>(2000);
    bad |= shifts<36>(2000) | shifts<48>(2000) | shifts<60>(2000) | shifts<64>(2000) | shifts<72>(2000) | shifts<84>(2000);
    printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}
