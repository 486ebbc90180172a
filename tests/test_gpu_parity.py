"""GPU parity tests (-m gpu): the CUDA path through the C ABI against the CPU oracle, bit-exact.

Cases follow the reference's own test shapes: the Poseidon KAT (poseidon256.rs:163-202), MerkleTree
new/prove/verify (poseidon_bn128.rs:209-267), and the commit shapes of SURVEY.md 3.4.
"""
import os

import numpy as np
import pytest

import oracle
import vectorx_b200 as vx
from oracle import pyref

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "commit_golden.npz"))
P = oracle.P


def test_round_constants_device_copy(ctx):
    rc = np.zeros(360, dtype=np.uint64)
    oracle.lib().vxo_poseidon_constants(oracle._ptr(rc))
    assert np.array_equal(vx.poseidon_round_constants(), rc)


def test_poseidon_golden_and_kat(ctx):
    assert np.array_equal(vx.poseidon(np.zeros((1, 12), dtype=np.uint64))[0], GOLD["perm_zero"])
    assert np.array_equal(vx.poseidon(np.arange(12, dtype=np.uint64)[None])[0], GOLD["perm_iota"])
    assert np.array_equal(vx.poseidon(GOLD["perm_in"]), GOLD["perm_out"])
    assert np.array_equal(vx.hash_n_to_hash_no_pad(GOLD["kat_in"][None])[0], GOLD["kat_out"])


def test_poseidon_random_and_noncanonical(ctx):
    rng = np.random.default_rng(5)
    st = rng.integers(0, 2**64, size=(4096, 12), dtype=np.uint64)          # includes values >= p
    st[0, :] = np.uint64(2**64 - 1); st[1, :] = np.uint64(P); st[2, :] = np.uint64(P - 1)
    st[3, :] = np.uint64(0xFFFFFFFF); st[4, :] = np.uint64(0xFFFFFFFF00000000)
    got = vx.poseidon(st)
    want = np.stack([oracle.poseidon(s) for s in st])
    assert np.array_equal(got, want)
    assert (got < np.uint64(P)).all()


@pytest.mark.parametrize("ln", [0, 1, 4, 5, 7, 8, 9, 16, 20, 32, 135])
def test_hash_no_pad_lengths(ctx, ln):
    assert np.array_equal(vx.hash_n_to_hash_no_pad(GOLD[f"hash_in_{ln}"][None])[0], GOLD[f"hash_out_{ln}"])
    x = oracle.random_field((300, ln), seed=900 + ln)
    want = np.stack([oracle.hash_no_pad(r) for r in x]) if ln else np.zeros((300, 4), dtype=np.uint64)
    assert np.array_equal(vx.hash_n_to_hash_no_pad(x), want)


@pytest.mark.parametrize("n,w,cap", [(1, 5, 0), (2, 9, 1), (16, 3, 4), (16, 4, 2), (64, 135, 4), (32, 20, 0),
                                     (8, 32, 3), (1 << 11, 32, 4), (1 << 12, 135, 4), (1 << 10, 1, 0), (256, 7, 8)])
def test_merkle_tree_new(ctx, n, w, cap):
    leaves = oracle.random_field((n, w), seed=n * 7 + w)
    digests, capv = vx.merkle_tree_digests(leaves, cap)
    od, oc = oracle.merkle_new(leaves, cap)
    assert np.array_equal(capv, oc)
    assert np.array_equal(digests, od)
    tree = vx.MerkleTree.new(leaves, cap)
    assert np.array_equal(tree.cap.hashes, oc)
    idx = sorted(set([0, n - 1, n // 2, n // 3]))
    proofs = tree.prove_many(idx)
    rows = tree.get_many(idx)
    for i, pr, row in zip(idx, proofs, rows):
        assert np.array_equal(row, leaves[i])
        assert np.array_equal(pr.siblings, oracle.merkle_prove(od, n, cap, i))
        assert oracle.merkle_verify(row, i, pr.siblings, capv)       # verify_merkle_proof_to_cap


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 8, 10, 12, 13, 16, 17, 20])
def test_ntt_forward_inverse(ctx, log_n):
    c = 3 if log_n < 18 else 1
    x = oracle.random_field((c, 1 << log_n), seed=log_n)
    y = vx.ntt(x)
    assert np.array_equal(y, np.stack([oracle.fft(r) for r in x]))
    assert np.array_equal(vx.ntt(y, inverse=True), x)
    if log_n <= 13:
        ys = vx.ntt(x, coset_shift=oracle.GENERATOR)
        assert np.array_equal(ys, np.stack([oracle.fft(r, shift=oracle.GENERATOR) for r in x]))
        assert np.array_equal(vx.ntt(ys, inverse=True, coset_shift=oracle.GENERATOR), x)


def _check_commit(cols, rate_bits, cap_height, from_coeffs=False):
    if from_coeffs:
        want = oracle.commit_from_coeffs(cols, rate_bits, cap_height)
        batch = vx.PolynomialBatch.from_coeffs(cols, rate_bits, False, cap_height)
    else:
        want = oracle.commit_from_values(cols, rate_bits, cap_height)
        batch = vx.PolynomialBatch.from_values(cols, rate_bits, False, cap_height)
    assert np.array_equal(batch.cap.hashes, want["cap"])
    assert np.array_equal(batch.polynomials, want["coeffs"])
    leaves, digests = batch.download()
    assert np.array_equal(leaves, want["leaves"])
    assert np.array_equal(digests, want["digests"])
    N = leaves.shape[0]
    idx = sorted(set([0, 1, N - 1, N // 2, (N * 5) // 7]))
    assert np.array_equal(batch.leaves(idx), want["leaves"][idx])
    paths = batch.prove(idx)
    for t, i in enumerate(idx):
        assert np.array_equal(paths[t], oracle.merkle_prove(want["digests"], N, cap_height, i))
    bits = batch.degree_log + rate_bits
    assert np.array_equal(batch.get_lde_values(3 % N, 1), want["leaves"][pyref.bitrev(3 % N, bits)])
    batch.close()
    return want


@pytest.mark.parametrize("c,log_n,rate,cap", [
    (3, 3, 1, 1),        # the hand-checkable tiny case (golden)
    (9, 1, 1, 2),        # cap == tree height (digests empty)
    (3, 6, 3, 4),        # narrow oracle: hash_or_noop leaves (<= 4 columns)
    (4, 5, 3, 4),
    (5, 5, 3, 4),        # first width that hashes
    (135, 10, 3, 4),     # standard_recursion_config wires, reduced height (golden)
    (20, 12, 3, 4),      # Z + partial products
    (16, 12, 3, 4),      # quotient chunks
    (85, 13, 3, 4),      # constants + sigmas
    (64, 14, 1, 4),      # starkyx rate_bits = 1
    (7, 16, 2, 0),       # cap_height 0, rate 2
    (1, 0, 3, 2),        # degenerate: one-row trace
])
def test_commit_from_values(ctx, c, log_n, rate, cap):
    cols = oracle.random_field((c, 1 << log_n), seed=c * 100 + log_n)
    _check_commit(cols, rate, cap)


def test_commit_from_coeffs(ctx):
    coeffs = oracle.random_field((16, 1 << 11), seed=77)
    _check_commit(coeffs, 3, 4, from_coeffs=True)


def test_commit_golden_fixtures(ctx):
    b = vx.PolynomialBatch.from_values(GOLD["tiny_cols"], 1, False, 1)
    leaves, digests = b.download()
    assert np.array_equal(leaves, GOLD["tiny_leaves"]) and np.array_equal(digests, GOLD["tiny_digests"])
    assert np.array_equal(b.cap.hashes, GOLD["tiny_cap"]) and np.array_equal(b.polynomials, GOLD["tiny_coeffs"])
    cols = oracle.random_field((135, 1 << 10), seed=int(GOLD["std_seed"][0]))
    b = vx.PolynomialBatch.from_values(cols, 3, False, 4)
    assert np.array_equal(b.cap.hashes, GOLD["std_cap"])
    assert np.array_equal(b.leaves(GOLD["std_idx"]), GOLD["std_rows"])
    assert np.array_equal(b.prove(GOLD["std_idx"]), GOLD["std_paths"])
    assert np.array_equal(b.polynomials[0], GOLD["std_coeffs_col0"])
    b = vx.PolynomialBatch.from_values(GOLD["narrow_cols"], 3, False, 4)
    assert np.array_equal(b.cap.hashes, GOLD["narrow_cap"])
    assert np.array_equal(b.download()[1], GOLD["narrow_digests"])
    b = vx.PolynomialBatch.from_values(GOLD["flat_cols"], 1, False, 2)
    assert np.array_equal(b.cap.hashes, GOLD["flat_cap"])


def test_structured_columns(ctx):
    """Real wire columns are mostly zeros / ones / small integers / non-canonical leftovers."""
    n = 1 << 10
    cols = np.zeros((6, n), dtype=np.uint64)
    cols[1, :] = 1
    cols[2, :] = np.arange(n, dtype=np.uint64)
    cols[3, ::2] = np.uint64(P - 1)
    cols[4, :] = np.uint64(2**64 - 1)          # non-canonical representative of 2^32 - 2
    cols[5, 5] = np.uint64(P)                  # non-canonical zero
    want = _check_commit(cols, 3, 4)
    assert (want["coeffs"][0] == 0).all() and int(want["coeffs"][1][0]) == 1


def test_full_size_properties_config1(ctx):
    """BASELINE config 1 (2^16 x 135, rate 3, cap 4): size-independent checks + sampled oracle parity."""
    c, log_n, rate, cap = 135, 16, 3, 4
    cols = oracle.random_field((c, 1 << log_n), seed=0x5EED0001)
    b = vx.PolynomialBatch.from_values(cols, rate, False, cap)
    bits = log_n + rate
    N = 1 << bits
    coeffs = b.polynomials
    # (1) iNTT parity on sampled columns
    for j in (0, 67, 134):
        assert np.array_equal(coeffs[j], oracle.fft(cols[j], inverse=True))
    # (2) LDE rows are Horner evaluations of the coefficients at g * w_N^i
    rng = np.random.default_rng(1)
    pts = [0, 1, N - 1] + [int(v) for v in rng.integers(0, N, size=5)]
    wN = pyref.primitive_root_of_unity(bits)
    rows = b.leaves([pyref.bitrev(i, bits) for i in pts])
    for t, i in enumerate(pts):
        x = oracle.GENERATOR * pow(wN, i, P) % P
        for j in (0, 1, 66, 134):
            assert int(rows[t][j]) == pyref.eval_poly([int(v) for v in coeffs[j]], x)
    # (3) every sampled Merkle path verifies against the cap under the oracle verifier
    idx = [0, N - 1] + [int(v) for v in rng.integers(0, N, size=30)]
    rows = b.leaves(idx)
    paths = b.prove(idx)
    capv = b.cap.hashes
    for t, i in enumerate(idx):
        assert oracle.merkle_verify(rows[t], i, paths[t], capv)
    # (4) the cap is the oracle's cap (full CPU commit, ~10 s)
    want = oracle.commit_from_values(cols, rate, cap, want_leaves=False, want_digests=False)
    assert np.array_equal(capv, want["cap"])
    # (5) linearity: commit(a) + commit(b) rows == commit(a + b) rows
    cols2 = oracle.random_field((c, 1 << log_n), seed=2)
    summed = ((cols.astype(object) + cols2.astype(object)) % P).astype(np.uint64)
    b2 = vx.PolynomialBatch.from_values(cols2, rate, False, cap)
    b3 = vx.PolynomialBatch.from_values(summed, rate, False, cap)
    r1, r2, r3 = b.leaves(idx[:6]), b2.leaves(idx[:6]), b3.leaves(idx[:6])
    assert np.array_equal(((r1.astype(object) + r2.astype(object)) % P).astype(np.uint64), r3)


@pytest.mark.parametrize("c,log_n,src", [(135, 9, "host"), (24, 10, "host"), (9, 8, "host"), (40, 9, "device")])
def test_commit_from_values_keep(ctx, c, log_n, src):
    """vx_commit_from_values_keep: the commitment equals from_values and the caller's device buffer holds the values
    (chunked streaming path for wide host inputs, single-copy path for narrow or device-resident ones)."""
    from vectorx_b200._lib import DeviceArray
    cols = oracle.random_field((c, 1 << log_n), seed=7 * c + log_n)
    want = oracle.commit_from_values(cols, 3, 4)
    keep = DeviceArray(ctx, cols.shape)
    source = cols if src == "host" else DeviceArray.from_host(ctx, cols)
    b = vx.PolynomialBatch.from_values_keep(source, keep, 3, 4, ctx=ctx)
    assert np.array_equal(b.cap.hashes, want["cap"])
    assert np.array_equal(b.polynomials, want["coeffs"])
    leaves, digests = b.download()
    assert np.array_equal(leaves, want["leaves"]) and np.array_equal(digests, want["digests"])
    assert np.array_equal(keep.to_host(), cols)
    b.close()
    keep.close()
    with pytest.raises(vx.VxError):
        vx.PolynomialBatch.from_values_keep(cols, cols.copy(), 3, 4, ctx=ctx)      # keep buffer must be device memory


@pytest.mark.parametrize("c,log_n,rate,cap", [(40, 13, 3, 1), (85, 15, 1, 3), (33, 14, 2, 0)])
def test_commit_device_resident_small_leaf_block(ctx, c, log_n, rate, cap):
    """A device-resident commit of 2^16 wide leaves: the shape of one rank's block of an 8-way sharded commit, hashed as two
    half-sponges per leaf block (leaf_hash_split_kernel, merkle.cu).  Digests, cap and leaves equal the oracle's."""
    from vectorx_b200._lib import DeviceArray
    cols = oracle.random_field((c, 1 << log_n), seed=1000 + c)
    want = oracle.commit_from_values(cols, rate, cap)
    dev = DeviceArray.from_host(ctx, cols)
    b = vx.PolynomialBatch.from_values(dev, rate, False, cap, ctx=ctx)
    assert np.array_equal(b.cap.hashes, want["cap"])
    leaves, digests = b.download()
    assert np.array_equal(leaves, want["leaves"]) and np.array_equal(digests, want["digests"])
    b.close()
    dev.close()


def test_pinned_host_buffers(ctx):
    """vx_host_alloc / vx_host_register: a commit whose values come from page-locked memory equals the pageable one."""
    from vectorx_b200._lib import check, load
    cols = oracle.random_field((9, 1 << 8), seed=31)
    want = oracle.commit_from_values(cols, 3, 4)
    pinned = vx.pinned_empty(cols.shape)
    assert pinned.shape == cols.shape and pinned.dtype == np.uint64 and pinned.flags["C_CONTIGUOUS"]
    pinned[:] = cols
    b = vx.PolynomialBatch.from_values(pinned, 3, False, 4)
    assert np.array_equal(b.cap.hashes, want["cap"])
    b.close()
    reg = cols.copy()
    check(load().vx_host_register(reg.ctypes.data, reg.nbytes), "vx_host_register")
    try:
        b = vx.PolynomialBatch.from_values(reg, 3, False, 4)
        assert np.array_equal(b.cap.hashes, want["cap"])
        b.close()
    finally:
        load().vx_host_unregister(reg.ctypes.data)
    assert load().vx_host_register(None, 0) != 0          # argument errors are codes, not crashes
    del pinned


def test_error_behaviour(ctx):
    cols = oracle.random_field((3, 6), seed=1)          # not a power of two
    with pytest.raises(vx.VxError):
        vx.PolynomialBatch.from_values(cols, 3, False, 4)
    cols = oracle.random_field((3, 8), seed=1)
    with pytest.raises(vx.VxError):
        vx.PolynomialBatch.from_values(cols, 1, False, 9)   # cap_height above tree height
    with pytest.raises(vx.VxError):
        vx.PolynomialBatch.from_values(cols, 1, True, 1)    # blinding unsupported
    b = vx.PolynomialBatch.from_values(cols, 1, False, 1)
    with pytest.raises(vx.VxError):
        b.leaves([16])                                      # out of range


@pytest.mark.parametrize("shards,c,log_n,rate", [(2, 21, 9, 3), (4, 21, 9, 3), (8, 21, 9, 3),
                                                 (16, 21, 9, 3),      # more shards than cosets: halves of a coset block
                                                 (4, 9, 13, 1), (8, 9, 14, 1), (16, 5, 16, 1),     # STARK rate: folded blocks
                                                 (8, 7, 10, 0), (4, 40, 12, 2)])
def test_sharded_commit_matches_whole(ctx, shards, c, log_n, rate):
    """SURVEY 8e coset partition: the shards' leaves / digests / caps concatenate to the 1-GPU commit -- whole cosets while
    shards <= 2^rate_bits, parts of one coset's leaf block (transform of the folded coefficients) beyond that."""
    cap = 4
    coeffs = oracle.random_field((c, 1 << log_n), seed=55)
    want = oracle.commit_from_coeffs(coeffs, rate, cap)
    N = (1 << log_n) << rate
    caps, leaves, digests = [], [], []
    for s in range(shards):
        b = vx.PolynomialBatch.from_coeffs_shard(coeffs, rate, cap, s, shards)
        assert (b.leaf_first, b.num_leaves, b.num_caps) == (s * N // shards, N // shards, 16 // shards)
        lv, dg = b.download()
        caps.append(b.cap.hashes); leaves.append(lv); digests.append(dg)
        i = 5
        assert np.array_equal(b.prove([i])[0], oracle.merkle_prove(want["digests"], N, cap, b.leaf_first + i))
    assert np.array_equal(np.concatenate(caps), want["cap"])
    assert np.array_equal(np.concatenate(leaves), want["leaves"])
    assert np.array_equal(np.concatenate(digests), want["digests"])


def test_field_primitives(ctx):
    """Goldilocks / quadratic-extension device arithmetic against Python big ints, incl. edge representatives."""
    from vectorx_b200._lib import check, load, ptr
    rng = np.random.default_rng(77)
    n = 4096
    edge = np.array([0, 1, 2, P - 1, P, P + 1, 2**64 - 1, 0xFFFFFFFF, 0x100000000, 0xFFFFFFFF00000000], dtype=np.uint64)
    a = rng.integers(0, 2**64, size=n, dtype=np.uint64); b = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    c = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    a[:10] = edge; b[:10] = edge[::-1]; a[10:20] = edge; b[10:20] = edge; c[:10] = edge
    A, B, C = [int(x) for x in a], [int(x) for x in b], [int(x) for x in c]

    def run(op, x, y=None, z=None, words=n):
        out = np.zeros(words, dtype=np.uint64)
        check(load().vx_field_op(ctx.handle, op, ptr(x), ptr(y) if y is not None else None,
                                 ptr(z) if z is not None else None, words if op < 7 else words // 2, ptr(out)), "vx_field_op")
        return [int(v) for v in out]
    assert run(0, a, b) == [x * y % P for x, y in zip(A, B)]
    assert run(1, a, b) == [x * y % P for x, y in zip(A, B)]
    assert run(3, a, b, c) == [(x * y + z) % P for x, y, z in zip(A, B, C)]
    assert run(4, a, b) == [(x + y) % P for x, y in zip(A, B)]
    assert run(5, a, b) == [(x - y) % P for x, y in zip(A, B)]
    assert run(6, a) == [pow(x, 7, P) for x in A]
    nz = np.where(a % np.uint64(P) == 0, np.uint64(5), a)
    assert run(2, nz) == [pow(int(x), P - 2, P) for x in nz]
    from oracle.field import E2
    got = run(7, a, b)
    for i in range(0, n, 2):
        w = E2(A[i], A[i + 1]) * E2(B[i], B[i + 1])
        assert got[i:i + 2] == [w.a, w.b]
    got = run(8, nz)
    for i in range(0, n, 2):
        w = E2(int(nz[i]), int(nz[i + 1])).inv()
        assert got[i:i + 2] == [w.a, w.b]
