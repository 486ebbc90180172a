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
