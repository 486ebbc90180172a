"""GPU parity tests (-m gpu) of the wrapper-stage hasher PoseidonBN128Hash: the CUDA path through the C ABI against the
reference's own known-answer test (P2X/backend/wrapper/poseidon_bn128.rs:134-181), the committed golden fixture and the
big-int oracle, bit-exact; tree cases follow poseidon_bn128.rs:205-267 (random_data(n, 7), log_n = 8, cap heights 1 and 8,
cap_height too big must fail)."""
import json
import os
import random

import numpy as np
import pytest

import oracle
import vectorx_b200 as vx
from oracle import bn128, pyref

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bn128_golden.json")))


def words(x: int):
    return bn128.to_limbs(x)


def state_words(s):
    return np.array([words(x) for x in s], dtype=np.uint64)


def test_permutation_reference_kat(ctx):
    st = np.stack([state_words(i) for i, _ in bn128.KAT])
    got = vx.poseidon_bn128(st, ctx=ctx)
    for k, (_, out) in enumerate(bn128.KAT):
        assert [bn128.from_limbs(got[k, i]) for i in range(4)] == out


def test_permutation_golden_and_random(ctx):
    st = np.stack([state_words([int(x) for x in i]) for i, _ in GOLD["permute"]])
    got = vx.poseidon_bn128(st, ctx=ctx)
    for k, (_, o) in enumerate(GOLD["permute"]):
        assert [bn128.from_limbs(got[k, i]) for i in range(4)] == [int(x) for x in o]
    rnd = random.Random(77)
    states = [[rnd.randrange(bn128.R) for _ in range(4)] for _ in range(200)]
    states += [[0, 0, 0, bn128.R - 1], [bn128.R - 1, 1, 0, 0], [1 << 253, (1 << 192) - 1, 1 << 64, 1 << 32]]
    got = vx.poseidon_bn128(np.stack([state_words(s) for s in states]), ctx=ctx)
    for k, s in enumerate(states):
        assert [bn128.from_limbs(got[k, i]) for i in range(4)] == bn128.permute(s), k


def test_permutation_rejects_non_canonical_scalars(ctx):
    st = state_words([0, 0, 0, bn128.R])[None]
    with pytest.raises(vx.VxError, match="modulus"):
        vx.poseidon_bn128(st, ctx=ctx)


def test_hash_golden(ctx):
    for key, noop in (("hash_no_pad", False), ("hash_or_noop", True)):
        for v, h in GOLD[key]:
            x = np.array([[int(e) for e in v]], dtype=np.uint64).reshape(1, len(v))
            got = vx.poseidon_bn128_hash(x, or_noop=noop, ctx=ctx)[0]
            assert bn128.from_limbs(got) == int(h), (key, len(v))


@pytest.mark.parametrize("ln", [1, 3, 4, 7, 9, 10, 20, 135])
def test_hash_random_batches_and_noncanonical_inputs(ctx, ln):
    rng = np.random.default_rng(ln)
    x = rng.integers(0, 2**64, size=(100, ln), dtype=np.uint64)           # includes representatives >= p
    x[0, :] = np.uint64(2**64 - 1)
    x[1, :] = np.uint64(oracle.P)
    for noop in (False, True):
        got = vx.poseidon_bn128_hash(x, or_noop=noop, ctx=ctx)
        f = bn128.hash_or_noop if noop else bn128.hash_no_pad
        for k in range(x.shape[0]):
            assert bn128.from_limbs(got[k]) == f([int(e) for e in x[k]]), (ln, noop, k)


@pytest.mark.parametrize("n,w,cap", [(256, 7, 1), (256, 7, 8), (16, 7, 1), (64, 3, 2), (32, 135, 0), (1, 10, 0), (128, 20, 4)])
def test_merkle_tree_new(ctx, n, w, cap):
    leaves = oracle.random_field((n, w), seed=n * 13 + w)
    digests, capv = vx.merkle_tree_digests(leaves, cap, ctx=ctx, hasher=vx.POSEIDON_BN128_HASH)
    od, oc = bn128.merkle_tree([[int(e) for e in r] for r in leaves], cap)
    assert [bn128.from_limbs(r) for r in capv] == oc
    assert [bn128.from_limbs(r) for r in digests] == od
    tree = vx.MerkleTree.new(leaves, cap, ctx=ctx, hasher=vx.POSEIDON_BN128_HASH)
    assert [bn128.from_limbs(r) for r in tree.cap.hashes] == oc
    idx = list(range(n)) if n <= 64 else [0, 1, n // 2 - 1, n // 2, n - 1, 77 % n]
    for i, proof in zip(idx, tree.prove_many(idx)):                       # verify_all_leaves, poseidon_bn128.rs:205-220
        sib = [bn128.from_limbs(s) for s in proof.siblings]
        assert sib == pyref.merkle_prove(od, n, cap, i)
        assert bn128.merkle_verify([int(e) for e in tree.get(i)], i, sib, oc)
    tree.close()


def test_golden_tree(ctx):
    t = GOLD["tree"]
    leaves = np.array([[int(x) for x in l] for l in t["leaves"]], dtype=np.uint64)
    digests, cap = vx.merkle_tree_digests(leaves, t["cap_height"], ctx=ctx, hasher=vx.POSEIDON_BN128_HASH)
    assert [str(bn128.from_limbs(r)) for r in digests] == t["digests"]
    assert [str(bn128.from_limbs(r)) for r in cap] == t["cap"]


def test_cap_height_too_big(ctx):
    """test_cap_height_too_big, poseidon_bn128.rs:222-233: cap_height = log_n + 1 must fail."""
    leaves = oracle.random_field((256, 7), seed=3)
    with pytest.raises(vx.VxError, match="cap_height"):
        vx.MerkleTree.new(leaves, 9, ctx=ctx, hasher=vx.POSEIDON_BN128_HASH)
    with pytest.raises(vx.VxError, match="hasher"):
        vx.MerkleTree.new(leaves, 1, ctx=ctx, hasher=7)


@pytest.mark.parametrize("log_n,c,rate_bits,cap", [(5, 10, 3, 4), (6, 3, 1, 0), (4, 135, 3, 2)])
def test_commit_with_bn128_hasher(ctx, log_n, c, rate_bits, cap):
    """PolynomialBatch::from_values under PoseidonBN128GoldilocksConfig (the wrap circuit's commitments): same LDE as
    the Goldilocks-Poseidon commit, the tree hashed with PoseidonBN128Hash."""
    cols = oracle.random_field((c, 1 << log_n), seed=log_n * 100 + c)
    want = oracle.commit_from_values(cols, rate_bits, cap)
    b = vx.PolynomialBatch.from_values(cols, rate_bits, False, cap, ctx=ctx, hasher=vx.POSEIDON_BN128_HASH)
    assert np.array_equal(b.polynomials, want["coeffs"])
    leaves, digests = b.download()
    assert np.array_equal(leaves, want["leaves"])
    od, oc = bn128.merkle_tree([[int(e) for e in r] for r in want["leaves"]], cap)
    assert [bn128.from_limbs(r) for r in b.cap.hashes] == oc
    assert [bn128.from_limbs(r) for r in digests] == od
    b2 = vx.PolynomialBatch.from_coeffs(want["coeffs"], rate_bits, False, cap, ctx=ctx, hasher=vx.POSEIDON_BN128_HASH)
    assert np.array_equal(b2.cap.hashes, b.cap.hashes)
    b.close(); b2.close()
