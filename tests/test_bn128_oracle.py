"""CPU tests of the wrapper-stage hasher (PoseidonBN128Hash): the big-int oracle against the reference's own known-answer
test and constants, the library's host-side table derivation against the oracle's, and the committed golden fixture.

Reference: P2X = contracts/lib/succinctx/plonky2x/core/src
  permutation KAT      P2X/backend/wrapper/poseidon_bn128.rs:134-181
  Merkle tree tests    P2X/backend/wrapper/poseidon_bn128.rs:205-267
  constants            P2X/backend/wrapper/poseidon_bn128_constants.rs (compared literal by literal when mounted)
"""
import importlib.util
import json
import os
import random

import pytest

import vectorx_b200 as vx
from oracle import bn128

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "bn128_golden.json")))


def _make_golden_module():
    spec = importlib.util.spec_from_file_location("make_bn128_golden", os.path.join(HERE, "golden", "make_bn128_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_kat_optimised_and_naive():
    for inp, out in bn128.KAT:
        assert bn128.permute(inp) == out
        assert bn128.permute_naive(inp) == out
    assert [[[int(x) for x in i], [int(x) for x in o]] for i, o in GOLD["kat"]] == [[i, o] for i, o in bn128.KAT]


def test_naive_and_optimised_schedules_agree():
    rnd = random.Random(1)
    for _ in range(10):
        s = [rnd.randrange(bn128.R) for _ in range(4)]
        assert bn128.permute(s) == bn128.permute_naive(s)


def test_tables_fingerprint_travels():
    assert bn128.tables_fingerprint() == GOLD["tables_sha256"]
    assert GOLD["tables_checked_against_reference_literals"] is True


def test_tables_equal_reference_literals():
    ref = _make_golden_module().reference_literals()
    if ref is None:
        pytest.skip("/root/reference not mounted (GPU box): the SHA-256 fingerprint test covers this")
    C, S, M, Pm = bn128.optimised_constants()
    assert ref[0] == C and ref[1] == S
    assert ref[2] == [x for r in M for x in r] and ref[3] == [x for r in Pm for x in r]


def test_library_host_derivation_matches_oracle():
    """vx_bn128_constants: C++ Grain LFSR + Montgomery arithmetic + sparse factorisation, no GPU involved."""
    C, S, M, Pm = vx.poseidon_bn128_constants()
    Co, So, Mo, Po = bn128.optimised_constants()
    assert C == Co and S == So
    assert M == [x for r in Mo for x in r] and Pm == [x for r in Po for x in r]


def test_golden_vectors():
    for i, o in GOLD["permute"]:
        assert bn128.permute([int(x) for x in i]) == [int(x) for x in o]
    for v, h in GOLD["hash_no_pad"]:
        assert bn128.hash_no_pad([int(x) for x in v]) == int(h)
    for v, h in GOLD["hash_or_noop"]:
        assert bn128.hash_or_noop([int(x) for x in v]) == int(h)
    for l, r, h in GOLD["two_to_one"]:
        assert bn128.two_to_one(int(l), int(r)) == int(h)


def test_hash_or_noop_is_identity_up_to_three_elements():
    v = [5, bn128.GL_P - 1, 7]
    assert bn128.hash_or_noop(v[:1]) == 5
    assert bn128.hash_or_noop(v) == 5 | ((bn128.GL_P - 1) << 64) | (7 << 128)
    assert bn128.hash_or_noop(v + [1]) == bn128.hash_no_pad(v + [1])
    assert bn128.hash_or_noop([bn128.GL_P + 5]) == 5                    # to_canonical_u64
    assert bn128.hash_pad([1, 2, 3]) == bn128.hash_no_pad([1, 2, 3, 1, 0, 0, 0, 0, 1])


def test_hash_out_to_vec_is_injective_chunks():
    h = bn128.R - 1
    v = bn128.hash_to_vec(h)
    assert len(v) == 5 and all(x < (1 << 56) for x in v)
    assert sum(x << (56 * i) for i, x in enumerate(v)) == h


@pytest.mark.parametrize("log_n,cap_height", [(4, 1), (4, 4), (3, 0)])
def test_merkle_trees_verify_all_leaves(log_n, cap_height):
    """test_merkle_trees / test_cap_height_eq_log2_len (poseidon_bn128.rs:235-267) at CPU-sized trees."""
    rnd = random.Random(log_n * 10 + cap_height)
    n = 1 << log_n
    leaves = [[rnd.randrange(bn128.GL_P) for _ in range(7)] for _ in range(n)]
    digests, cap = bn128.merkle_tree(leaves, cap_height)
    sub = n >> cap_height
    from oracle import pyref
    for i, leaf in enumerate(leaves):
        sib = pyref.merkle_prove(digests, n, cap_height, i)
        assert len(sib) == sub.bit_length() - 1
        assert bn128.merkle_verify(leaf, i, sib, cap)
        bad = list(leaf)
        bad[0] = (bad[0] + 1) % bn128.GL_P
        assert not bn128.merkle_verify(bad, i, sib, cap)


def test_cap_height_too_big():
    with pytest.raises(AssertionError):
        bn128.merkle_tree([[1] * 7] * 8, 4)


def test_golden_tree():
    t = GOLD["tree"]
    leaves = [[int(x) for x in l] for l in t["leaves"]]
    dg, cap = bn128.merkle_tree(leaves, t["cap_height"])
    assert [str(x) for x in dg] == t["digests"] and [str(x) for x in cap] == t["cap"]
