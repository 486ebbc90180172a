"""GPU parity at the shapes BASELINE.json's config 5 sweep and the STARK traces name, above the 2^16 rows
the fast tests stop at: the strided NTT pass changes its grid tiling there (grid.y x grid.z), the four-step tables change
size and, past 2^20 rows, the passes fall back to the two-level twiddle tables.

 * full CPU cap comparison (native oracle, all host cores) at 2^18 x 64 rate 3, 2^20 x 128 rate 2, 2^18 x 1271 rate 1;
 * 2^22 and 2^24 rows: values that are the evaluations of SPARSE polynomials (forward NTT on the CPU oracle), so that the
   coefficients the commit must return are known exactly, every LDE row is a cheap closed-form evaluation at g w_N^i, and
   sampled Merkle paths verify against the cap;
 * the per-column entry point (plonky2's Vec<PolynomialValues>: separately allocated, pageable columns) and concurrent
   commits from several threads on one context (SURVEY.md 8b threading contract).
"""
import threading

import numpy as np
import pytest

import oracle
import vectorx_b200 as vx
from oracle import P, pyref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c,log_n,rate", [(64, 18, 3), (128, 20, 2), (1271, 18, 1)])
def test_cap_equals_full_cpu_commit(ctx, c, log_n, rate):
    cap = 4
    cols = oracle.random_field((c, 1 << log_n), seed=1000 * log_n + c)
    b = vx.PolynomialBatch.from_values(cols, rate, False, cap, ctx=ctx)
    want = oracle.commit_from_values(cols, rate, cap, want_leaves=False, want_digests=False, native=True)
    assert np.array_equal(b.cap.hashes, want["cap"])
    # sampled coefficient columns against the CPU iNTT, sampled Merkle paths against the cap
    coeffs = b.polynomials
    for j in (0, c // 2, c - 1):
        assert np.array_equal(coeffs[j], want["coeffs"][j])
    del coeffs
    N = 1 << (log_n + rate)
    rng = np.random.default_rng(5)
    idx = [0, N - 1] + [int(v) for v in rng.integers(0, N, size=10)]
    rows, paths = b.leaves(idx), b.prove(idx)
    for t, i in enumerate(idx):
        assert oracle.merkle_verify(rows[t], i, paths[t], b.cap.hashes)
    b.close()


def _sparse(c, n, seed):
    """c sparse coefficient columns: {exponent: value} with a constant term, two random terms and the top coefficient"""
    rng = np.random.default_rng(seed)
    out = []
    for j in range(c):
        terms = {0: int(rng.integers(1, P, dtype=np.uint64)), n - 1: int(rng.integers(1, P, dtype=np.uint64))}
        for e in rng.integers(1, n - 1, size=2):
            terms[int(e)] = int(rng.integers(1, P, dtype=np.uint64))
        out.append(terms)
    return out


@pytest.mark.parametrize("log_n,c,rate", [(22, 6, 2), (24, 5, 1)])
def test_sparse_polynomials_at_sweep_heights(ctx, log_n, c, rate):
    n, cap = 1 << log_n, 4
    terms = _sparse(c, n, seed=log_n)
    coeffs = np.zeros((c, n), dtype=np.uint64)
    for j, t in enumerate(terms):
        for e, v in t.items():
            coeffs[j, e] = v
    vals = np.stack([oracle.fft(coeffs[j]) for j in range(c)])           # CPU forward NTT: the subgroup evaluations
    b = vx.PolynomialBatch.from_values(vals, rate, False, cap, ctx=ctx)
    assert np.array_equal(b.polynomials, coeffs)
    bits = log_n + rate
    N = 1 << bits
    wN = pyref.primitive_root_of_unity(bits)
    rng = np.random.default_rng(9)
    pts = [0, 1, N - 1, N // 2 + 1] + [int(v) for v in rng.integers(0, N, size=12)]
    leaf_idx = [pyref.bitrev(i, bits) for i in pts]
    rows, paths = b.leaves(leaf_idx), b.prove(leaf_idx)
    for t, i in enumerate(pts):
        x = oracle.GENERATOR * pow(wN, i, P) % P
        for j in range(c):
            want = sum(v * pow(x, e, P) for e, v in terms[j].items()) % P
            assert int(rows[t][j]) == want, (i, j)
        assert oracle.merkle_verify(rows[t], leaf_idx[t], paths[t], b.cap.hashes)
    # the same through from_coeffs: identical cap
    b2 = vx.PolynomialBatch.from_coeffs(coeffs, rate, False, cap, ctx=ctx)
    assert np.array_equal(b2.cap.hashes, b.cap.hashes)
    b.close(); b2.close()


@pytest.mark.parametrize("c,log_n,rate", [(135, 12, 3), (20, 13, 3), (64, 14, 1), (5, 9, 2), (16, 16, 3)])
def test_commit_from_separately_allocated_columns(ctx, c, log_n, rate):
    """vx_commit_from_values_cols / _coeffs_cols: c separate pageable allocations (plonky2's Vec<PolynomialValues>)."""
    n, cap = 1 << log_n, 4
    flat = oracle.random_field((c, n), seed=31 * c + log_n)
    cols = [flat[j].copy() for j in range(c)]                          # c independent heap allocations
    want = vx.PolynomialBatch.from_values(flat, rate, False, cap, ctx=ctx)
    got = vx.PolynomialBatch.from_values(cols, rate, False, cap, ctx=ctx)
    assert np.array_equal(got.cap.hashes, want.cap.hashes)
    assert np.array_equal(got.polynomials, want.polynomials)
    if log_n <= 13:
        ref = oracle.commit_from_values(flat, rate, cap)
        leaves, digests = got.download()
        assert np.array_equal(got.cap.hashes, ref["cap"])
        assert np.array_equal(leaves, ref["leaves"]) and np.array_equal(digests, ref["digests"])
    coeff_cols = [want.polynomials[j].copy() for j in range(c)]
    got2 = vx.PolynomialBatch.from_coeffs(coeff_cols, rate, False, cap, ctx=ctx)
    assert np.array_equal(got2.cap.hashes, want.cap.hashes)
    # device-resident columns (torch tensors) go through the same entry point
    import torch
    dev_cols = [torch.from_numpy(x.view(np.int64)).to("cuda:0") for x in cols]
    got3 = vx.PolynomialBatch.from_values(dev_cols, rate, False, cap, ctx=ctx)
    assert np.array_equal(got3.cap.hashes, want.cap.hashes)
    for b in (want, got, got2, got3):
        b.close()


def test_concurrent_commits_on_one_context(ctx):
    """Several host threads commit on ONE context at the same time (Rayon workers, several STARK proofs in flight):
    every call takes its own lane (stream set) and every result equals the single-threaded one."""
    shapes = [(40, 12, 3), (135, 11, 3), (17, 13, 1), (64, 12, 2)]
    inputs, want = [], []
    for t, (c, log_n, rate) in enumerate(shapes):
        cols = oracle.random_field((c, 1 << log_n), seed=500 + t)
        inputs.append(cols)
        want.append(oracle.commit_from_values(cols, rate, 4, want_leaves=False, want_digests=False)["cap"])
    errors = []

    def worker(t):
        try:
            c, log_n, rate = shapes[t]
            for rep in range(6):
                src = inputs[t] if rep % 2 == 0 else [inputs[t][j].copy() for j in range(c)]
                b = vx.PolynomialBatch.from_values(src, rate, False, 4, ctx=ctx)
                if not np.array_equal(b.cap.hashes, want[t]):
                    errors.append(f"thread {t} rep {rep}: cap mismatch")
                idx = [0, 5, (1 << (log_n + rate)) - 1]
                rows, paths = b.leaves(idx), b.prove(idx)
                for r, i in enumerate(idx):
                    if not oracle.merkle_verify(rows[r], i, paths[r], b.cap.hashes):
                        errors.append(f"thread {t} rep {rep}: path {i} does not verify")
                b.close()
        except Exception as e:       # noqa: BLE001
            errors.append(f"thread {t}: {e!r}")

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(len(shapes))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
