"""Big-int Python restatement of the plonky2 v0.2.0 commit path (TEST INFRASTRUCTURE ONLY).

This file is the *second, independent* oracle: plain Python integers, naive
algorithms, small sizes only.  It exists to cross-check the C oracle (oracle/oracle.c) and is pinned
by the one golden vector the reference holds at this boundary:

    /root/reference/contracts/lib/succinctx/plonky2x/core/src/frontend/hash/poseidon/poseidon256.rs:163-202

The algorithms themselves live in third-party crates that are NOT vendored under /root/reference:
plonky2 + plonky2_field v0.2.0 @ 7445ec9 (Cargo.lock:4847-4850, 4871-4873).  They are restated from
their published algorithm; reference call sites are cited per function.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

P = 0xFFFFFFFF00000001
EPS = 0xFFFFFFFF
GENERATOR = 14293326489335486720          # MULTIPLICATIVE_GROUP_GENERATOR == coset_shift()
POWER_OF_TWO_GENERATOR = 7277203076849721926   # order 2^32
TWO_ADICITY = 32
W_EXT = 7                                  # F[x]/(x^2 - 7)

MASK64 = (1 << 64) - 1
MASK32 = (1 << 32) - 1


# --------------------------------------------------------------------------- field
def inv(a: int) -> int:
    return pow(a % P, P - 2, P)


def primitive_root_of_unity(k: int) -> int:
    """plonky2_field: POWER_OF_TWO_GENERATOR ^ (2^(32-k))."""
    assert 0 <= k <= TWO_ADICITY
    return pow(POWER_OF_TWO_GENERATOR, 1 << (TWO_ADICITY - k), P)


def bitrev(i: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


# --------------------------------------------------------------------------- ChaCha8Rng (SURVEY 8c recipe)
def _pcg32_seed(seed_u64: int) -> list[int]:
    """rand_core::SeedableRng::seed_from_u64: 8 PCG32 outputs -> 32-byte key (as 8 LE u32 words)."""
    MUL, INC = 6364136223846793005, 11634580027462260723
    state, words = seed_u64, []
    for _ in range(8):
        state = (state * MUL + INC) & MASK64
        xorshifted = (((state >> 18) ^ state) >> 27) & MASK32
        rot = state >> 59
        words.append(((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & MASK32)
    return words


def _rotl(x, n):
    return ((x << n) | (x >> (32 - n))) & MASK32


def _chacha_block(key: list[int], counter: int, rounds: int = 8) -> list[int]:
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + key + [
        counter & MASK32, (counter >> 32) & MASK32, 0, 0]
    x = st[:]

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & MASK32; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & MASK32; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & MASK32; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & MASK32; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + st[i]) & MASK32 for i in range(16)]


class ChaCha8Rng:
    def __init__(self, seed_u64: int):
        self.key = _pcg32_seed(seed_u64)
        self.counter = 0
        self.buf: list[int] = []

    def next_u32(self) -> int:
        if not self.buf:
            self.buf = _chacha_block(self.key, self.counter)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)

    def gen_range_p(self) -> int:
        """rand 0.8 UniformInt::<u64>::sample_single(0, P): widening multiply + zone rejection."""
        zone = P - 1            # (P << P.leading_zeros()) - 1, leading_zeros == 0
        while True:
            v = self.next_u64()
            m = v * P
            if (m & MASK64) <= zone:
                return m >> 64


_RC = None


def round_constants() -> list[int]:
    """360 = 30 rounds x 12 lanes: plonky2 hash/poseidon_goldilocks.rs ALL_ROUND_CONSTANTS."""
    global _RC
    if _RC is None:
        rng = ChaCha8Rng(0)
        _RC = [rng.gen_range_p() for _ in range(360)]
    return _RC


# --------------------------------------------------------------------------- Poseidon
MDS_CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
MDS_DIAG = [8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
HALF_N_FULL_ROUNDS = 4
N_PARTIAL_ROUNDS = 22
WIDTH = 12
RATE = 8


def mds(state: list[int]) -> list[int]:
    return [(sum(state[(i + r) % 12] * MDS_CIRC[i] for i in range(12)) + state[r] * MDS_DIAG[r]) % P
            for r in range(12)]


def poseidon(state: list[int]) -> list[int]:
    """plonky2 hash/poseidon.rs Poseidon::poseidon in its naive (spec) form."""
    rc = round_constants()
    s = [x % P for x in state]
    for r in range(30):
        s = [(s[i] + rc[12 * r + i]) % P for i in range(12)]
        if r < 4 or r >= 26:
            s = [pow(x, 7, P) for x in s]
        else:
            s[0] = pow(s[0], 7, P)
        s = mds(s)
    return s


def hash_n_to_hash_no_pad(inputs: list[int]) -> list[int]:
    """plonky2 hash/hashing.rs; also called at P2X/utils/poseidon/mod.rs:31-36. Overwrite-mode sponge."""
    state = [0] * 12
    for off in range(0, len(inputs), RATE):
        chunk = inputs[off:off + RATE]
        state[:len(chunk)] = [x % P for x in chunk]
        state = poseidon(state)
    return state[:4]


def two_to_one(left: list[int], right: list[int]) -> list[int]:
    return poseidon(list(left) + list(right) + [0, 0, 0, 0])[:4]


def hash_or_noop(inputs: list[int]) -> list[int]:
    """Hasher::hash_or_noop (shape shown in-tree at P2X/backend/wrapper/plonky2_config.rs:130-196)."""
    if len(inputs) <= 4:
        return [x % P for x in inputs] + [0] * (4 - len(inputs))
    return hash_n_to_hash_no_pad(inputs)


# --------------------------------------------------------------------------- Merkle
def merkle_tree(leaves: list[list[int]], cap_height: int):
    """Returns (digests, cap) in plonky2's interleaved layout."""
    n = len(leaves)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n and cap_height <= log_n
    num_caps = 1 << cap_height
    sub = n >> cap_height
    digests: list = [None] * (2 * (n - num_caps))
    cap = []
    for s in range(num_caps):
        base = s * (2 * sub - 2)
        layer = [hash_or_noop(leaves[s * sub + j]) for j in range(sub)]
        lvl = 0
        while len(layer) > 1:
            for q in range(len(layer) // 2):
                pos = base + 2 * (q * (1 << (lvl + 1)) + (1 << lvl) - 1)
                digests[pos] = layer[2 * q]
                digests[pos + 1] = layer[2 * q + 1]
            layer = [two_to_one(layer[2 * q], layer[2 * q + 1]) for q in range(len(layer) // 2)]
            lvl += 1
        cap.append(layer[0])
    return digests, cap


def merkle_prove(digests, n_leaves: int, cap_height: int, leaf_index: int) -> list[list[int]]:
    sub = n_leaves >> cap_height
    s, j = divmod(leaf_index, sub)
    base = s * (2 * sub - 2)
    sib = []
    lvl = 0
    while (sub >> lvl) > 1:
        idx = j >> lvl
        q = idx >> 1
        pos = base + 2 * (q * (1 << (lvl + 1)) + (1 << lvl) - 1)
        sib.append(digests[pos + ((idx & 1) ^ 1)])
        lvl += 1
    return sib


def merkle_verify(leaf: list[int], leaf_index: int, siblings, cap) -> bool:
    cur = hash_or_noop(leaf)
    idx = leaf_index
    for sib in siblings:
        cur = two_to_one(sib, cur) if idx & 1 else two_to_one(cur, sib)
        idx >>= 1
    return cur == cap[idx]


# --------------------------------------------------------------------------- NTT, naive
def fft(coeffs: list[int]) -> list[int]:
    """out[k] = sum_j c_j w^{jk}; recursive radix-2 (small sizes)."""
    n = len(coeffs)
    if n == 1:
        return [coeffs[0] % P]
    log_n = n.bit_length() - 1
    w = primitive_root_of_unity(log_n)
    ev = fft(coeffs[0::2])
    od = fft(coeffs[1::2])
    out = [0] * n
    t = 1
    for k in range(n // 2):
        x = od[k] * t % P
        out[k] = (ev[k] + x) % P
        out[k + n // 2] = (ev[k] - x) % P
        t = t * w % P
    return out


def ifft(values: list[int]) -> list[int]:
    n = len(values)
    buf = fft(values)
    ninv = inv(n)
    return [buf[(n - i) % n] * ninv % P for i in range(n)]


def coset_fft(coeffs: list[int], shift: int) -> list[int]:
    t, scaled = 1, []
    for c in coeffs:
        scaled.append(c * t % P)
        t = t * shift % P
    return fft(scaled)


def lde_values(coeffs: list[int], rate_bits: int) -> list[int]:
    """PolynomialBatch::lde_values for one column: zero-pad then coset_fft(g)."""
    n = len(coeffs)
    return coset_fft(list(coeffs) + [0] * (n * ((1 << rate_bits) - 1)), GENERATOR)


def commit_from_coeffs(coeff_cols: list[list[int]], rate_bits: int, cap_height: int):
    """fri/oracle.rs PolynomialBatch::from_coeffs -> (leaves in bit-reversed row order, digests, cap)."""
    n = len(coeff_cols[0])
    big_n = n << rate_bits
    bits = big_n.bit_length() - 1
    ldes = [lde_values(c, rate_bits) for c in coeff_cols]
    leaves = [[ldes[j][bitrev(i, bits)] for j in range(len(coeff_cols))] for i in range(big_n)]
    digests, cap = merkle_tree(leaves, cap_height)
    return leaves, digests, cap


def commit_from_values(value_cols: list[list[int]], rate_bits: int, cap_height: int):
    coeffs = [ifft(v) for v in value_cols]
    leaves, digests, cap = commit_from_coeffs(coeffs, rate_bits, cap_height)
    return coeffs, leaves, digests, cap


def eval_poly(coeffs: list[int], x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


# --------------------------------------------------------------------------- the reference KAT
def kat_poseidon256():
    """test_poseidon, P2X/frontend/hash/poseidon/poseidon256.rs:163-202 (encoding: :96-128, vars/byte.rs:49-57)."""
    leaf = bytes.fromhex("d68d62c262c2ec08961c1104188cde86f51695878759666ad61490c8ec66745c")
    expected = bytes.fromhex("faa1095f1959da5713d6ad8b21b54936f167dc8e3f205b129b8eb8740aa10c0b")
    bits = []
    for b in leaf:                      # ByteVariable holds bits MSB first
        bits += [(b >> (7 - i)) & 1 for i in range(8)]
    inputs = [sum(bit << k for k, bit in enumerate(bits[32 * w:32 * w + 32])) for w in range(8)]  # le_sum
    out = hash_n_to_hash_no_pad(inputs)
    got = bytearray()
    for e in out:
        le_bits = [(e >> k) & 1 for k in range(64)]      # split_le
        for m in range(8):
            got.append(sum(le_bits[8 * m + i] << (7 - i) for i in range(8)))
    return inputs, out, bytes(got), expected


if __name__ == "__main__":
    rc = round_constants()
    print([hex(x) for x in rc[:4]], hex(rc[-1]))
    print([hex(x) for x in poseidon([0] * 12)[:4]])
    inputs, out, got, exp = kat_poseidon256()
    print(got.hex(), exp.hex(), got == exp)
