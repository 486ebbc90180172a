"""CPU tests: the oracle against the reference KAT, the committed golden vectors and itself.

The oracle is test infrastructure (oracle/oracle.h).  These tests pin it to the only golden vector
the reference holds at this boundary (poseidon256.rs:163-202), cross-check the C restatement with the
independent big-int Python restatement, and hold the rest by algebraic self-checks (end).
"""
import os
import random

import numpy as np
import pytest

import oracle
from oracle import pyref

P = oracle.P
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "commit_golden.npz"))


def test_reference_kat_poseidon256():
    # contracts/lib/succinctx/plonky2x/core/src/frontend/hash/poseidon/poseidon256.rs:172-178
    inputs, out, got, expected = pyref.kat_poseidon256()
    assert got == expected
    assert [int(x) for x in oracle.hash_no_pad(inputs)] == out


def test_round_constants_match_known_values():
    rc = np.zeros(360, dtype=np.uint64)
    oracle.lib().vxo_poseidon_constants(oracle._ptr(rc))
    assert [int(x) for x in rc] == pyref.round_constants()
    # first four entries of plonky2's ALL_ROUND_CONSTANTS
    assert [hex(int(x)) for x in rc[:4]] == ["0xb585f766f2144405", "0x7746a55f43921ad7",
                                               "0xb2fb0d31cee799b4", "0xf6760a4803427d7"]
    assert all(int(x) < P for x in rc)


def test_permutation_of_zero():
    got = [hex(int(x)) for x in oracle.poseidon([0] * 12)[:4]]
    assert got == ["0x3c18a9786cb0b359", "0xc4055e3364a246c3", "0x7953db0ab48808f4", "0xc71603f33a1144ca"]


def test_c_poseidon_vs_bigint_and_naive():
    rnd = random.Random(7)
    for _ in range(8):
        st = [rnd.randrange(2**64) for _ in range(12)]       # non-canonical inputs allowed
        want = pyref.poseidon(st)
        assert [int(x) for x in oracle.poseidon(st)] == want
        s2 = np.array(st, dtype=np.uint64)
        oracle.lib().vxo_poseidon_naive(oracle._ptr(s2))
        assert [int(x) for x in s2] == want


@pytest.mark.parametrize("ln", [0, 1, 4, 5, 7, 8, 9, 16, 20, 32, 135])
def test_sponge_lengths(ln):
    x = GOLD[f"hash_in_{ln}"]
    assert np.array_equal(oracle.hash_no_pad(x), GOLD[f"hash_out_{ln}"])
    assert [int(v) for v in oracle.hash_no_pad(x)] == pyref.hash_n_to_hash_no_pad([int(v) for v in x])


def test_field_identities():
    L = oracle.lib()
    rnd = random.Random(3)
    for _ in range(200):
        a, b = rnd.randrange(2**64), rnd.randrange(2**64)
        assert L.vxo_mul(a, b) == (a * b) % P
        assert L.vxo_add(a, b) == (a + b) % P
        assert L.vxo_sub(a, b) == (a - b) % P
    for a in (1, 2, P - 1, 0xFFFFFFFF, 0xFFFFFFFF00000000):
        assert L.vxo_mul(a, L.vxo_inv(a)) == 1
    # two-adic structure
    assert pow(oracle.GENERATOR, (P - 1) // 2**32, P) == 7277203076849721926
    assert L.vxo_root_of_unity(6) == 8 and L.vxo_root_of_unity(5) == 64 and L.vxo_root_of_unity(4) == 4096
    e = np.array([3, 5], dtype=np.uint64); f = np.array([7, 11], dtype=np.uint64); o = np.zeros(2, dtype=np.uint64)
    L.vxo_ext_mul(oracle._ptr(e), oracle._ptr(f), oracle._ptr(o))
    assert o.tolist() == [3 * 7 + 7 * 5 * 11, 3 * 11 + 5 * 7]
    L.vxo_ext_inv(oracle._ptr(e), oracle._ptr(o))
    L.vxo_ext_mul(oracle._ptr(e), oracle._ptr(o), oracle._ptr(f))
    assert f.tolist() == [1, 0]


@pytest.mark.parametrize("log_n", [0, 1, 3, 8, 12])
def test_fft_roundtrip_and_definition(log_n):
    n = 1 << log_n
    x = oracle.random_field((n,), seed=log_n)
    y = oracle.fft(x)
    assert np.array_equal(oracle.fft(y, inverse=True), x)
    if n <= 8:
        w = pyref.primitive_root_of_unity(log_n)
        want = [sum(int(x[j]) * pow(w, j * k, P) for j in range(n)) % P for k in range(n)]
        assert [int(v) for v in y] == want
    ys = oracle.fft(x, shift=oracle.GENERATOR)
    assert np.array_equal(oracle.fft(ys, inverse=True, shift=oracle.GENERATOR), x)


def test_tiny_commit_golden_and_bigint():
    cols = GOLD["tiny_cols"]
    r = oracle.commit_from_values(cols, 1, 1)
    for k in ("coeffs", "leaves", "digests", "cap"):
        assert np.array_equal(r[k], GOLD[f"tiny_{k}"]), k
    pc, pl, pd, pcap = pyref.commit_from_values([[int(v) for v in c] for c in cols], 1, 1)
    assert r["coeffs"].tolist() == pc and r["leaves"].tolist() == pl
    assert r["digests"].tolist() == pd and r["cap"].tolist() == pcap


def test_commit_is_evaluation_of_coefficients():
    """Horner self-check: leaves[bitrev(i)][col] == coeffs_col(g * w_N^i) (check 1)."""
    c, log_n, rate = 5, 6, 3
    cols = oracle.random_field((c, 1 << log_n), seed=11)
    r = oracle.commit_from_values(cols, rate, 2)
    bits = log_n + rate
    wN = pyref.primitive_root_of_unity(bits)
    for i in (0, 1, 7, 100, (1 << bits) - 1):
        x = oracle.GENERATOR * pow(wN, i, P) % P
        row = r["leaves"][pyref.bitrev(i, bits)]
        for j in range(c):
            assert int(row[j]) == pyref.eval_poly([int(v) for v in r["coeffs"][j]], x)
    # and the coefficients interpolate the input values on the subgroup
    wn = pyref.primitive_root_of_unity(log_n)
    for k in (0, 3, 63):
        assert pyref.eval_poly([int(v) for v in r["coeffs"][2]], pow(wn, k, P)) == int(cols[2][k])


@pytest.mark.parametrize("n,w,cap", [(1, 5, 0), (2, 9, 1), (16, 3, 4), (16, 4, 2), (64, 135, 4), (32, 20, 0), (8, 32, 3)])
def test_merkle_layout_prove_verify(n, w, cap):
    leaves = oracle.random_field((n, w), seed=n * 131 + w)
    digests, capv = oracle.merkle_new(leaves, cap)
    pd, pcap = pyref.merkle_tree([[int(v) for v in row] for row in leaves], cap)
    assert digests.tolist() == pd and capv.tolist() == pcap
    for j in range(n):
        sib = oracle.merkle_prove(digests, n, cap, j)
        assert sib.tolist() == pyref.merkle_prove(pd, n, cap, j)
        assert oracle.merkle_verify(leaves[j], j, sib, capv)
        if sib.size:
            bad = sib.copy(); bad[0, 0] ^= np.uint64(1)
            assert not oracle.merkle_verify(leaves[j], j, bad, capv)


def test_std_shape_golden():
    seed = int(GOLD["std_seed"][0])
    cols = oracle.random_field((135, 1 << 10), seed=seed)
    r = oracle.commit_from_values(cols, 3, 4)
    assert np.array_equal(r["cap"], GOLD["std_cap"])
    idx = GOLD["std_idx"].astype(np.int64)
    assert np.array_equal(r["leaves"][idx], GOLD["std_rows"])
    assert np.array_equal(r["digests"][:64], GOLD["std_digests_head"])
    for t, i in enumerate(idx):
        assert oracle.merkle_verify(GOLD["std_rows"][t], int(i), GOLD["std_paths"][t], GOLD["std_cap"])
