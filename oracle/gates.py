"""Oracle-side gate constraint formulas (TEST INFRASTRUCTURE).

Each gate's `eval(w, c, pi)` returns the list of constraint values in plonky2's yield order, written
with +, -, * only so it runs on FA (all LDE points at once), FI and E2 (verifier at zeta).

Sources.  In-tree (exact): U32ArithmeticGate, U32SubtractionGate, U32RangeCheckGate --
contracts/lib/succinctx/plonky2x/core/src/frontend/uint/num/u32/gates/{arithmetic_u32.rs:290-349,
subtraction_u32.rs:235-271, range_check_u32.rs:93-115}.  Upstream plonky2 v0.2.0 gates (not vendored;
restated per SURVEY.md Appendix B): NoopGate, ConstantGate, PublicInputGate, ArithmeticGate,
BaseSumGate<2>, PoseidonGate.  The registry the reference serialises is
.../backend/circuit/serialization/gates.rs:85-107.
"""
from __future__ import annotations

from . import pyref

NUM_WIRES = 135
NUM_ROUTED = 80


class Gate:
    name = "?"
    degree = 0
    num_constants = 0
    num_constraints = 0

    def eval(self, w, c, pi):
        raise NotImplementedError

    def id(self):
        return self.name


class NoopGate(Gate):
    name = "NoopGate"
    degree = 0

    def eval(self, w, c, pi):
        return []


class ConstantGate(Gate):
    def __init__(self, num_consts=2):
        self.num_consts = num_consts
        self.name = f"ConstantGate {{ num_consts: {num_consts} }}"
        self.degree = 1
        self.num_constants = num_consts
        self.num_constraints = num_consts

    def eval(self, w, c, pi):
        return [c[i] - w[i] for i in range(self.num_consts)]


class PublicInputGate(Gate):
    name = "PublicInputGate"
    degree = 1
    num_constraints = 4

    def eval(self, w, c, pi):
        return [w[i] - pi[i] for i in range(4)]


class ArithmeticGate(Gate):
    def __init__(self, num_ops=20):
        self.num_ops = num_ops
        self.name = f"ArithmeticGate {{ num_ops: {num_ops} }}"
        self.degree = 3
        self.num_constants = 2
        self.num_constraints = num_ops

    def eval(self, w, c, pi):
        out = []
        for i in range(self.num_ops):
            m0, m1, add, res = w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]
            out.append(res - (m0 * m1 * c[0] + add * c[1]))
        return out


class BaseSumGate(Gate):
    """BaseSumGate<2>: wire 0 = sum, wires 1..=num_limbs = bits (little endian)."""

    def __init__(self, num_limbs=63):
        self.num_limbs = num_limbs
        self.name = f"BaseSumGate {{ num_limbs: {num_limbs} }} + Base: 2"
        self.degree = 2
        self.num_constraints = 1 + num_limbs

    def eval(self, w, c, pi):
        limbs = [w[1 + i] for i in range(self.num_limbs)]
        acc = limbs[-1]
        for l in reversed(limbs[:-1]):          # reduce_with_powers(limbs, 2), Horner
            acc = acc * 2 + l
        out = [acc - w[0]]
        for l in limbs:
            out.append(l * (l - 1))
        return out


class U32ArithmeticGate(Gate):
    """arithmetic_u32.rs:290-349; num_ops = min(135/38, 80/6) = 3."""

    def __init__(self, num_ops=3):
        self.num_ops = num_ops
        self.name = f"U32ArithmeticGate {{ num_ops: {num_ops} }}"
        self.degree = 4
        self.num_constraints = num_ops * (3 + 32 + 1)

    def eval(self, w, c, pi):
        out = []
        n = self.num_ops
        for i in range(n):
            m0, m1, add = w[6 * i], w[6 * i + 1], w[6 * i + 2]
            lo, hi, inv = w[6 * i + 3], w[6 * i + 4], w[6 * i + 5]
            computed = m0 * m1 + add
            diff = (0xFFFFFFFF - hi)
            hi_not_max = inv * diff - 1
            out.append(hi_not_max * lo)
            out.append(hi * (1 << 32) + lo - computed)
            comb_lo = comb_hi = None
            for j in reversed(range(32)):
                limb = w[6 * n + 32 * i + j]
                out.append(limb * (limb - 1) * (limb - 2) * (limb - 3))
                if j < 16:
                    comb_lo = limb if comb_lo is None else comb_lo * 4 + limb
                else:
                    comb_hi = limb if comb_hi is None else comb_hi * 4 + limb
            out.append(comb_lo - lo)
            out.append(comb_hi - hi)
        return out


class U32SubtractionGate(Gate):
    """subtraction_u32.rs:235-271; num_ops = min(135/21, 80/5) = 6; 16 two-bit limbs per op."""

    def __init__(self, num_ops=6):
        self.num_ops = num_ops
        self.name = f"U32SubtractionGate {{ num_ops: {num_ops} }}"
        self.degree = 4
        self.num_constraints = num_ops * (1 + 16 + 1 + 1)

    def eval(self, w, c, pi):
        out = []
        n = self.num_ops
        for i in range(n):
            x, y, bin_, res, bout = (w[5 * i + k] for k in range(5))
            out.append(res - (x - y - bin_ + bout * (1 << 32)))
            comb = None
            for j in reversed(range(16)):
                limb = w[5 * n + 16 * i + j]
                out.append(limb * (limb - 1) * (limb - 2) * (limb - 3))
                comb = limb if comb is None else comb * 4 + limb
            out.append(comb - res)
            out.append(bout * (1 - bout))
        return out


class U32RangeCheckGate(Gate):
    """range_check_u32.rs:93-115: inputs 0..k, 16 two-bit aux limbs per input after them."""

    def __init__(self, num_input_limbs=7):
        self.k = num_input_limbs
        self.name = f"U32RangeCheckGate {{ num_input_limbs: {num_input_limbs} }}"
        self.degree = 4
        self.num_constraints = num_input_limbs * (1 + 16)

    def eval(self, w, c, pi):
        out = []
        k = self.k
        for i in range(k):
            aux = [w[k + 16 * i + j] for j in range(16)]
            comb = None
            for l in reversed(aux):
                comb = l if comb is None else comb * 4 + l
            out.append(comb - w[i])
            for l in aux:
                out.append(l * (l - 1) * (l - 2) * (l - 3))
        return out


class PoseidonGate(Gate):
    """plonky2 gates/poseidon.rs: one width-12 permutation per row, S-box inputs wired (degree 7)."""
    name = "PoseidonGate"
    degree = 7
    num_constraints = 123
    W_IN, W_OUT, W_SWAP, W_DELTA = 0, 12, 24, 25
    W_FULL0, W_PARTIAL, W_FULL1 = 29, 65, 87

    def eval(self, w, c, pi):
        rc = pyref.round_constants()
        out = []
        swap = w[self.W_SWAP]
        out.append(swap * (swap - 1))
        delta = [w[self.W_DELTA + i] for i in range(4)]
        for i in range(4):
            out.append(swap * (w[i + 4] - w[i]) - delta[i])
        st = [None] * 12
        for i in range(4):
            st[i] = w[i] + delta[i]
            st[i + 4] = w[i + 4] - delta[i]
        for i in range(8, 12):
            st[i] = w[i]

        def mds(s):
            return [sum((s[(i + r) % 12] * pyref.MDS_CIRC[i] for i in range(1, 12)), s[r] * pyref.MDS_CIRC[0])
                    + (s[r] * pyref.MDS_DIAG[r] if pyref.MDS_DIAG[r] else 0) for r in range(12)]

        def sbox(x):
            x2 = x * x
            x4 = x2 * x2
            return x4 * x2 * x

        rnd = 0
        for r in range(4):
            st = [st[i] + rc[12 * rnd + i] for i in range(12)]
            if r != 0:
                for i in range(12):
                    wi = w[self.W_FULL0 + 12 * (r - 1) + i]
                    out.append(st[i] - wi)
                    st[i] = wi
            st = mds([sbox(x) for x in st])
            rnd += 1
        for r in range(22):                      # naive partial rounds: same lane-0 S-box inputs as the fast form
            st = [st[i] + rc[12 * rnd + i] for i in range(12)]
            wi = w[self.W_PARTIAL + r]
            out.append(st[0] - wi)
            st[0] = sbox(wi)
            st = mds(st)
            rnd += 1
        for r in range(4):
            st = [st[i] + rc[12 * rnd + i] for i in range(12)]
            for i in range(12):
                wi = w[self.W_FULL1 + 12 * r + i]
                out.append(st[i] - wi)
                st[i] = wi
            st = mds([sbox(x) for x in st])
            rnd += 1
        for i in range(12):
            out.append(st[i] - w[self.W_OUT + i])
        assert len(out) == 123
        return out

    @staticmethod
    def fill_witness(row, inputs, swap=0):
        """Sets wires of `row` (a list of 135 ints) for a permutation of `inputs` (12 ints)."""
        P = pyref.P
        rc = pyref.round_constants()
        G = PoseidonGate
        for i in range(12):
            row[G.W_IN + i] = inputs[i] % P
        row[G.W_SWAP] = swap
        st = [x % P for x in inputs]
        for i in range(4):
            d = swap * (st[i + 4] - st[i]) % P
            row[G.W_DELTA + i] = d
            st[i], st[i + 4] = (st[i] + d) % P, (st[i + 4] - d) % P
        rnd = 0
        for r in range(4):
            st = [(st[i] + rc[12 * rnd + i]) % P for i in range(12)]
            if r != 0:
                for i in range(12):
                    row[G.W_FULL0 + 12 * (r - 1) + i] = st[i]
            st = pyref.mds([pow(x, 7, P) for x in st])
            rnd += 1
        for r in range(22):
            st = [(st[i] + rc[12 * rnd + i]) % P for i in range(12)]
            row[G.W_PARTIAL + r] = st[0]
            st[0] = pow(st[0], 7, P)
            st = pyref.mds(st)
            rnd += 1
        for r in range(4):
            st = [(st[i] + rc[12 * rnd + i]) % P for i in range(12)]
            for i in range(12):
                row[G.W_FULL1 + 12 * r + i] = st[i]
            st = pyref.mds([pow(x, 7, P) for x in st])
            rnd += 1
        for i in range(12):
            row[G.W_OUT + i] = st[i]
        return st


# ------------------------------------------------------------------------------------------------ in-tree gates (exact)
class U32AddManyGate(Gate):
    """add_many_u32.rs:41-84 (layout), :149-190 (constraints): per op `num_addends` addends + carry -> 32-bit result +
    output carry, decomposed into 16 + 3 two-bit limbs."""
    NUM_RESULT_LIMBS, NUM_CARRY_LIMBS = 16, 3

    def __init__(self, num_addends=3, num_ops=None):
        a = num_addends
        self.num_addends = a
        self.num_ops = num_ops if num_ops is not None else min(NUM_WIRES // (a + 3 + 19), NUM_ROUTED // (a + 3))
        self.name = f"U32AddManyGate {{ num_addends: {a}, num_ops: {self.num_ops} }}"
        self.degree = 4
        self.num_constraints = self.num_ops * (3 + 19)

    def w_addend(self, i, j): return (self.num_addends + 3) * i + j
    def w_carry(self, i): return (self.num_addends + 3) * i + self.num_addends
    def w_result(self, i): return (self.num_addends + 3) * i + self.num_addends + 1
    def w_out_carry(self, i): return (self.num_addends + 3) * i + self.num_addends + 2
    def w_limb(self, i, j): return (self.num_addends + 3) * self.num_ops + 19 * i + j

    def eval(self, w, c, pi):
        out = []
        for i in range(self.num_ops):
            computed = w[self.w_carry(i)]
            for j in range(self.num_addends):
                computed = computed + w[self.w_addend(i, j)]
            res, oc = w[self.w_result(i)], w[self.w_out_carry(i)]
            out.append(oc * (1 << 32) + res - computed)
            comb_res = comb_carry = None
            for j in reversed(range(19)):
                limb = w[self.w_limb(i, j)]
                out.append(limb * (limb - 1) * (limb - 2) * (limb - 3))
                if j < 16:
                    comb_res = limb if comb_res is None else comb_res * 4 + limb
                else:
                    comb_carry = limb if comb_carry is None else comb_carry * 4 + limb
            out.append(comb_res - res)
            out.append(comb_carry - oc)
        return out

    def fill_witness(self, row, rnd):
        for i in range(self.num_ops):
            total = 0
            for j in range(self.num_addends):
                v = rnd.randrange(1 << 32)
                row[self.w_addend(i, j)] = v
                total += v
            carry = rnd.randrange(1 << 32)
            row[self.w_carry(i)] = carry
            total += carry
            res, oc = total & 0xFFFFFFFF, total >> 32
            row[self.w_result(i)], row[self.w_out_carry(i)] = res, oc
            for j in range(16):
                row[self.w_limb(i, j)] = (res >> (2 * j)) & 3
            for j in range(3):
                row[self.w_limb(i, 16 + j)] = (oc >> (2 * j)) & 3


class ComparisonGate(Gate):
    """comparison.rs:44-92 (layout), :333-410 (constraints): result = (first <= second) on num_bits-bit inputs split
    into num_chunks chunks (gadget use: multiple_comparison.rs:39, 32 bits in 16 chunks)."""

    def __init__(self, num_bits=32, num_chunks=16):
        self.num_bits, self.num_chunks = num_bits, num_chunks
        self.chunk_bits = -(-num_bits // num_chunks)
        self.name = f"ComparisonGate {{ num_bits: {num_bits}, num_chunks: {num_chunks} }}<D=2>"
        self.degree = 1 << self.chunk_bits
        self.num_constraints = 6 + 5 * num_chunks + self.chunk_bits

    def eval(self, w, c, pi):
        n, cb = self.num_chunks, self.chunk_bits
        first = [w[4 + i] for i in range(n)]
        second = [w[4 + n + i] for i in range(n)]

        def horner(xs, base):
            acc = xs[-1]
            for x in reversed(xs[:-1]):
                acc = acc * base + x
            return acc

        out = [horner(first, 1 << cb) - w[0], horner(second, 1 << cb) - w[1]]
        msd = None
        for i in range(n):
            p1 = p2 = None
            for x in range(1 << cb):
                t1, t2 = first[i] - x, second[i] - x
                p1 = t1 if p1 is None else p1 * t1
                p2 = t2 if p2 is None else p2 * t2
            out += [p1, p2]
            diff = second[i] - first[i]
            dummy, eq, inter = w[4 + 2 * n + i], w[4 + 3 * n + i], w[4 + 4 * n + i]
            out.append(diff * dummy - (1 - eq))
            out.append(eq * diff)
            out.append(inter - eq * msd if msd is not None else inter - eq * 0)
            msd = inter + (1 - eq) * diff
        out.append(w[3] - msd)
        bits = [w[4 + 5 * n + i] for i in range(cb + 1)]
        for b in bits:
            out.append(b * (1 - b))
        out.append(w[3] + (1 << cb) - horner(bits, 2))
        out.append(w[2] - bits[cb])
        assert len(out) == self.num_constraints
        return out

    def fill_witness(self, row, rnd, a=None, b=None):
        P = pyref.P
        n, cb = self.num_chunks, self.chunk_bits
        a = rnd.randrange(1 << self.num_bits) if a is None else a
        b = rnd.randrange(1 << self.num_bits) if b is None else b
        row[0], row[1] = a, b
        msd = 0
        for i in range(n):
            x, y = (a >> (cb * i)) & ((1 << cb) - 1), (b >> (cb * i)) & ((1 << cb) - 1)
            row[4 + i], row[4 + n + i] = x, y
            diff = (y - x) % P
            eq = 1 if x == y else 0
            row[4 + 2 * n + i] = pow(diff, P - 2, P) if diff else 1        # equality dummy
            row[4 + 3 * n + i] = eq
            row[4 + 4 * n + i] = eq * msd % P                              # intermediate value
            msd = (row[4 + 4 * n + i] + (1 - eq) * diff) % P
        row[3] = msd
        v = ((1 << cb) + (msd if msd < P // 2 else msd - P))               # 2^cb + signed msd in [1, 2^(cb+1))
        for i in range(cb + 1):
            row[4 + 5 * n + i] = (v >> i) & 1
        row[2] = (v >> cb) & 1
        assert row[2] == (1 if a <= b else 0)


# ------------------------------------------------------------------------------------------------ upstream gates
# plonky2 v0.2.0 gates restated from SURVEY.md Appendix B (source not vendored in /root/reference): PARITY UNPINNED;
# each is validated by an honest-witness test (constraints vanish) and by GPU == oracle on random wires.
D = 2
W7 = 7          # F[x]/(x^2 - 7)


def ext_mul(a, b):
    return (a[0] * b[0] + a[1] * b[1] * W7, a[0] * b[1] + a[1] * b[0])


def _ext_mul_int(a, b):
    P = pyref.P
    return ((a[0] * b[0] + W7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


class ArithmeticExtensionGate(Gate):
    """gates/arithmetic_extension.rs: per op out = m0 * m1 * c0 + addend * c1 over the quadratic extension."""

    def __init__(self, num_ops=NUM_ROUTED // (4 * D)):
        self.num_ops = num_ops
        self.name = f"ArithmeticExtensionGate {{ num_ops: {num_ops} }}"
        self.degree, self.num_constants, self.num_constraints = 3, 2, num_ops * D

    def eval(self, w, c, pi):
        out = []
        for i in range(self.num_ops):
            b = 4 * D * i
            m0, m1, add, res = (w[b], w[b + 1]), (w[b + 2], w[b + 3]), (w[b + 4], w[b + 5]), (w[b + 6], w[b + 7])
            pr = ext_mul(m0, m1)
            for k in range(D):
                out.append(res[k] - (pr[k] * c[0] + add[k] * c[1]))
        return out

    def fill_witness(self, row, rnd, c0, c1):
        P = pyref.P
        for i in range(self.num_ops):
            b = 4 * D * i
            v = [rnd.randrange(P) for _ in range(6)]
            pr = _ext_mul_int((v[0], v[1]), (v[2], v[3]))
            row[b:b + 6] = v
            row[b + 6] = (pr[0] * c0 + v[4] * c1) % P
            row[b + 7] = (pr[1] * c0 + v[5] * c1) % P


class MulExtensionGate(Gate):
    """gates/multiplication_extension.rs: per op out = m0 * m1 * c0."""

    def __init__(self, num_ops=NUM_ROUTED // (3 * D)):
        self.num_ops = num_ops
        self.name = f"MulExtensionGate {{ num_ops: {num_ops} }}"
        self.degree, self.num_constants, self.num_constraints = 3, 1, num_ops * D

    def eval(self, w, c, pi):
        out = []
        for i in range(self.num_ops):
            b = 3 * D * i
            pr = ext_mul((w[b], w[b + 1]), (w[b + 2], w[b + 3]))
            for k in range(D):
                out.append(w[b + 4 + k] - pr[k] * c[0])
        return out

    def fill_witness(self, row, rnd, c0):
        P = pyref.P
        for i in range(self.num_ops):
            b = 3 * D * i
            v = [rnd.randrange(P) for _ in range(4)]
            pr = _ext_mul_int((v[0], v[1]), (v[2], v[3]))
            row[b:b + 4] = v
            row[b + 4], row[b + 5] = pr[0] * c0 % P, pr[1] * c0 % P


class ReducingGate(Gate):
    """gates/reducing.rs: Horner accumulation acc_i = acc_{i-1} * alpha + coeff_i with base-field coefficients;
    wires: output 0..D, alpha D..2D, old_acc 2D..3D, coeffs from 3D, intermediate accumulators after them (the last
    accumulator is the output)."""
    EXT_COEFFS = False

    def __init__(self, num_coeffs=None):
        if num_coeffs is None:
            num_coeffs = (NUM_WIRES - 2 * D) // (D + (D if self.EXT_COEFFS else 1))
            num_coeffs = min(num_coeffs, (NUM_ROUTED - 3 * D) // (D if self.EXT_COEFFS else 1))
        self.num_coeffs = num_coeffs
        self.name = f"{type(self).__name__} {{ num_coeffs: {num_coeffs} }}"
        self.degree, self.num_constants, self.num_constraints = 2, 0, D * num_coeffs

    def w_coeff(self, i):
        return (3 * D + D * i, 3 * D + D * i + 1) if self.EXT_COEFFS else (3 * D + i, None)

    def w_acc(self, i):
        if i == self.num_coeffs - 1:
            return 0
        return 3 * D + self.num_coeffs * (D if self.EXT_COEFFS else 1) + D * i

    def eval(self, w, c, pi):
        alpha = (w[D], w[D + 1])
        acc = (w[2 * D], w[2 * D + 1])
        out = []
        for i in range(self.num_coeffs):
            c0, c1 = self.w_coeff(i)
            a = self.w_acc(i)
            pr = ext_mul(acc, alpha)
            out.append(pr[0] + w[c0] - w[a])
            out.append(pr[1] + w[c1] - w[a + 1] if c1 is not None else pr[1] - w[a + 1])
            acc = (w[a], w[a + 1])
        return out

    def fill_witness(self, row, rnd):
        P = pyref.P
        alpha = (rnd.randrange(P), rnd.randrange(P))
        acc = (rnd.randrange(P), rnd.randrange(P))
        row[D], row[D + 1] = alpha
        row[2 * D], row[2 * D + 1] = acc
        for i in range(self.num_coeffs):
            c0, c1 = self.w_coeff(i)
            row[c0] = rnd.randrange(P)
            if c1 is not None:
                row[c1] = rnd.randrange(P)
            pr = _ext_mul_int(acc, alpha)
            acc = ((pr[0] + row[c0]) % P, (pr[1] + (row[c1] if c1 is not None else 0)) % P)
            a = self.w_acc(i)
            row[a], row[a + 1] = acc


class ReducingExtensionGate(ReducingGate):
    """gates/reducing_extension.rs: the same with extension-field coefficients."""
    EXT_COEFFS = True


class ExponentiationGate(Gate):
    """gates/exponentiation.rs: output = base^(power bits), square-and-multiply from the top bit with every
    intermediate value wired; wires: base 0, bits 1..1+n (little endian), output 1+n, intermediates from 2+n."""

    def __init__(self, num_power_bits=(NUM_WIRES - 2) // 2):
        n = num_power_bits
        self.n = n
        self.name = f"ExponentiationGate {{ num_power_bits: {n} }}<D=2>"
        self.degree, self.num_constants, self.num_constraints = 4, 0, n + 1

    def eval(self, w, c, pi):
        n = self.n
        out = []
        for i in range(n):
            bit = w[1 + (n - 1 - i)]
            factor = bit * w[0] + (1 - bit)
            if i == 0:
                computed = factor
            else:
                prev = w[2 + n + i - 1]
                computed = prev * prev * factor
            out.append(computed - w[2 + n + i])
        out.append(w[1 + n] - w[2 + n + n - 1])
        return out

    def fill_witness(self, row, rnd):
        P = pyref.P
        n = self.n
        base, e = rnd.randrange(P), rnd.randrange(1 << n)
        row[0] = base
        cur = 1
        for i in range(n):
            row[1 + i] = (e >> i) & 1
        for i in range(n):
            bit = (e >> (n - 1 - i)) & 1
            cur = cur * cur % P * (base if bit else 1) % P
            row[2 + n + i] = cur
        row[1 + n] = cur
        assert cur == pow(base, e, P)


class PoseidonMdsGate(Gate):
    """gates/poseidon_mds.rs: 12 extension inputs (wires 0..24) -> 12 extension outputs (24..48), out = MDS * in."""
    name = "PoseidonMdsGate"
    degree, num_constants, num_constraints = 1, 0, 12 * D

    def eval(self, w, c, pi):
        out = []
        for r in range(12):
            for k in range(D):
                acc = w[D * r + k] * (pyref.MDS_CIRC[0] + pyref.MDS_DIAG[r])
                for i in range(1, 12):
                    acc = acc + w[D * ((i + r) % 12) + k] * pyref.MDS_CIRC[i]
                out.append(w[D * (12 + r) + k] - acc)
        return out

    def fill_witness(self, row, rnd):
        P = pyref.P
        for i in range(24):
            row[i] = rnd.randrange(P)
        for k in range(D):
            res = pyref.mds([row[D * i + k] for i in range(12)])
            for r in range(12):
                row[D * (12 + r) + k] = res[r]


class RandomAccessGate(Gate):
    """gates/random_access.rs: per copy claimed = list[index] for a 2^bits list; index bits are unrouted wires after the
    routed ones; `num_extra_constants` constant wires ride along."""

    def __init__(self, bits=4, num_copies=None, num_extra_constants=None):
        vs = 1 << bits
        if num_copies is None:
            num_copies = min(NUM_ROUTED // (2 + vs), NUM_WIRES // (2 + vs + bits))
        if num_extra_constants is None:
            num_extra_constants = min(2, NUM_ROUTED - (2 + vs) * num_copies)
        self.bits, self.num_copies, self.num_extra = bits, num_copies, num_extra_constants
        self.vs = vs
        self.name = (f"RandomAccessGate {{ bits: {bits}, num_copies: {num_copies}, "
                     f"num_extra_constants: {num_extra_constants} }}<D=2>")
        self.degree, self.num_constants = bits + 1, num_extra_constants
        self.num_constraints = num_copies * (bits + 2) + num_extra_constants

    def num_routed(self): return (2 + self.vs) * self.num_copies + self.num_extra
    def w_bit(self, i, copy): return self.num_routed() + copy * self.bits + i

    def eval(self, w, c, pi):
        out = []
        for cp in range(self.num_copies):
            b0 = (2 + self.vs) * cp
            bits = [w[self.w_bit(i, cp)] for i in range(self.bits)]
            for b in bits:
                out.append(b * (b - 1))
            rec = None
            for b in reversed(bits):
                rec = b if rec is None else rec * 2 + b
            out.append(rec - w[b0])
            items = [w[b0 + 2 + i] for i in range(self.vs)]
            for b in bits:
                items = [items[2 * k] + b * (items[2 * k + 1] - items[2 * k]) for k in range(len(items) // 2)]
            out.append(items[0] - w[b0 + 1])
        for i in range(self.num_extra):
            out.append(c[i] - w[(2 + self.vs) * self.num_copies + i])
        return out

    def fill_witness(self, row, rnd, consts):
        P = pyref.P
        for cp in range(self.num_copies):
            b0 = (2 + self.vs) * cp
            idx = rnd.randrange(self.vs)
            items = [rnd.randrange(P) for _ in range(self.vs)]
            row[b0], row[b0 + 1] = idx, items[idx]
            row[b0 + 2:b0 + 2 + self.vs] = items
            for i in range(self.bits):
                row[self.w_bit(i, cp)] = (idx >> i) & 1
        for i in range(self.num_extra):
            row[(2 + self.vs) * self.num_copies + i] = consts[i]


def coset_interpolation_shape(subgroup_bits: int, max_degree: int):
    """gates/coset_interpolation.rs `with_max_degree`: (degree, num_intermediates) for a 2^subgroup_bits-point coset."""
    n_points = 1 << subgroup_bits
    n_intermediates = (n_points - 2) // (max_degree - 1)
    degree = (n_points - 2) // (n_intermediates + 1) + 2
    return degree, (n_points - 2) // (degree - 1)


def coset_interpolation_tables(subgroup_bits: int):
    """(domain, barycentric weights): domain = two_adic_subgroup in natural order, w_i = 1 / prod_{j != i} (x_i - x_j)."""
    P = pyref.P
    g = pyref.primitive_root_of_unity(subgroup_bits)
    dom = [pow(g, i, P) for i in range(1 << subgroup_bits)]
    weights = []
    for i, xi in enumerate(dom):
        d = 1
        for j, xj in enumerate(dom):
            if j != i:
                d = d * (xi - xj) % P
        weights.append(pow(d, P - 2, P))
    return dom, weights


class CosetInterpolationGate(Gate):
    """gates/coset_interpolation.rs (plonky2 v0.2.0; restated -- the upstream source is not in the reference tree, so wire
    order and constraint order are anchored only by the honest-witness and prove -> verify self-checks): barycentric
    interpolation of 2^subgroup_bits extension values given on the coset shift*<w>, evaluated at an extension point.
    Wires: shift 0; values 1 + D i; evaluation point, evaluation value after them (end of the routed wires); then
    `num_intermediates` partial evaluations, as many partial products, and the shifted evaluation point point/shift.
    Constraints: point - shifted*shift; per intermediate (eval, prod) against the fold over the next degree-1 points;
    value - final eval.  Fold step: eval' = eval (x - x_i) + w_i v_i prod, prod' = prod (x - x_i)."""

    def __init__(self, subgroup_bits=4, max_degree=8):
        self.bits = subgroup_bits
        self.np = 1 << subgroup_bits
        self.deg, self.ni = coset_interpolation_shape(subgroup_bits, max_degree)
        self.domain, self.weights = coset_interpolation_tables(subgroup_bits)
        self.name = (f"CosetInterpolationGate {{ subgroup_bits: {subgroup_bits}, degree: {self.deg}, "
                     f"barycentric_weights: {self.weights}, _phantom: PhantomData<plonky2_field::goldilocks_field::"
                     f"GoldilocksField> }}<D=2>")
        self.degree, self.num_constants = self.deg, 0
        self.num_constraints = D * (2 + 2 * self.ni)

    def w_value(self, i): return 1 + D * i
    def w_point(self): return 1 + D * self.np
    def w_eval_value(self): return self.w_point() + D
    def start_intermediates(self): return self.w_eval_value() + D
    def num_routed(self): return self.start_intermediates()
    def w_inter_eval(self, i): return self.start_intermediates() + D * i
    def w_inter_prod(self, i): return self.start_intermediates() + D * (self.ni + i)
    def w_shifted(self): return self.start_intermediates() + D * 2 * self.ni
    def end(self): return self.start_intermediates() + D * (2 * self.ni + 1)

    def chunks(self):
        """Point ranges folded between consecutive (intermediate) constraints."""
        out = [(0, self.deg)]
        for i in range(self.ni):
            s = 1 + (self.deg - 1) * (i + 1)
            out.append((s, min(s + self.deg - 1, self.np)))
        return out

    def _fold(self, w, lo, hi, x, ev, pr):
        for i in range(lo, hi):
            term = (x[0] - self.domain[i], x[1])
            v = (w[self.w_value(i)] * self.weights[i], w[self.w_value(i) + 1] * self.weights[i])
            vp = ext_mul(v, pr)
            et = ext_mul(ev, term)
            ev = (et[0] + vp[0], et[1] + vp[1])
            pr = ext_mul(pr, term)
        return ev, pr

    def eval(self, w, c, pi):
        shift = w[0]
        pt, sh = self.w_point(), self.w_shifted()
        x = (w[sh], w[sh + 1])
        out = [w[pt] - x[0] * shift, w[pt + 1] - x[1] * shift]
        ch = self.chunks()
        zero, one = w[0] * 0, w[0] * 0 + 1
        ev, pr = self._fold(w, ch[0][0], ch[0][1], x, (zero, zero), (one, zero))
        for i in range(self.ni):
            ie, ip = self.w_inter_eval(i), self.w_inter_prod(i)
            out += [w[ie] - ev[0], w[ie + 1] - ev[1], w[ip] - pr[0], w[ip + 1] - pr[1]]
            ev, pr = self._fold(w, ch[i + 1][0], ch[i + 1][1], x, (w[ie], w[ie + 1]), (w[ip], w[ip + 1]))
        e = self.w_eval_value()
        out += [w[e] - ev[0], w[e + 1] - ev[1]]
        return out

    def fill_witness(self, row, rnd):
        P = pyref.P
        shift = rnd.randrange(1, P)
        row[0] = shift
        for i in range(D * self.np):
            row[1 + i] = rnd.randrange(P)
        point = (rnd.randrange(P), rnd.randrange(P))
        inv = pow(shift, P - 2, P)
        x = (point[0] * inv % P, point[1] * inv % P)
        row[self.w_point()], row[self.w_point() + 1] = point
        row[self.w_shifted()], row[self.w_shifted() + 1] = x

        def fold(lo, hi, ev, pr):
            for i in range(lo, hi):
                term = ((x[0] - self.domain[i]) % P, x[1])
                v = (row[self.w_value(i)] * self.weights[i] % P, row[self.w_value(i) + 1] * self.weights[i] % P)
                vp, et = _ext_mul_int(v, pr), _ext_mul_int(ev, term)
                ev = ((et[0] + vp[0]) % P, (et[1] + vp[1]) % P)
                pr = _ext_mul_int(pr, term)
            return ev, pr
        ch = self.chunks()
        ev, pr = fold(ch[0][0], ch[0][1], (0, 0), (1, 0))
        for i in range(self.ni):
            row[self.w_inter_eval(i)], row[self.w_inter_eval(i) + 1] = ev
            row[self.w_inter_prod(i)], row[self.w_inter_prod(i) + 1] = pr
            ev, pr = fold(ch[i + 1][0], ch[i + 1][1], ev, pr)
        row[self.w_eval_value()], row[self.w_eval_value() + 1] = ev
        return x, ev

    def interpolate_direct(self, row, x):
        """Lagrange interpolation of the wired values at x (independent of the fold): the value the gate must output."""
        P = pyref.P
        acc = (0, 0)
        for i in range(self.np):
            num = (1, 0)
            den = 1
            for j in range(self.np):
                if j != i:
                    num = _ext_mul_int(num, ((x[0] - self.domain[j]) % P, x[1]))
                    den = den * (self.domain[i] - self.domain[j]) % P
            li = pow(den, P - 2, P)
            t = _ext_mul_int((row[self.w_value(i)], row[self.w_value(i) + 1]), num)
            acc = ((acc[0] + t[0] * li) % P, (acc[1] + t[1] * li) % P)
        return acc
