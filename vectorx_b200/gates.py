"""Gate constraint programs for the device interpreter (host side of vx_quotient).

What the Rust shim would obtain by tracing `Gate::eval_unfiltered_circuit` for every gate in
`CommonCircuitData.gates` (registry: contracts/lib/succinctx/plonky2x/core/src/backend/circuit/
serialization/gates.rs:85-107) is built here by tracing Python statements of the same formulas over
symbolic values: arithmetic on `Sym` objects records SSA ops, which are register-allocated into the
bytecode of include/vectorx_b200.h.

Formulas: U32* gates and ComparisonGate follow the in-tree sources (frontend/uint/num/u32/gates/
arithmetic_u32.rs:290-349, subtraction_u32.rs:235-271, range_check_u32.rs:93-115, add_many_u32.rs:149-190,
comparison.rs:333-410); the upstream plonky2 v0.2.0 gates follow SURVEY.md Appendix B.  19 of the 23 registered
gate types have a program (CosetInterpolationGate restated from plonky2 v0.2.0 without the source at hand: anchored by the
honest-witness, direct-Lagrange and prove -> verify checks only); LookupGate / LookupTableGate never occur (no lookup
tables in VectorX), ArithmeticCubicGate and MulCubicGate (starkyx) need source that is not in the reference tree.  PoseidonGate uses the fast partial-round tables exported by the library
(vx_poseidon_fast_tables), like upstream's gate does.
"""
from __future__ import annotations

import numpy as np

from ._lib import check, load, ptr

P = 0xFFFFFFFF00000001
OP = dict(END=0, LOADW=1, LOADC=2, LOADPI=3, LOADK=4, ADD=5, SUB=6, MUL=7, ADDK=8, MULK=9, RSUBK=10, SUBK=11,
          EMIT=12, BEGINGATE=13, ENDGATE=14, NOP=15, SBOX7=16, MDS12K=17, DENSE12=18, PARTIAL12=19, RANGE4=20, MADK=21)
MULTI = ("MDS12K", "DENSE12", "PARTIAL12")        # 12 registers in, 12 registers out
NUM_REGS = 64
LOAD_HOIST = 8          # column loads emitted together (one memory latency per group on the device)
LOAD_WINDOW = 400       # how far ahead (in traced operations) a load may be pulled
LOAD_RESERVE = 40       # registers that must stay free for the rest of the gate while hoisting
MDS_CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
MDS_DIAG0 = 8


class Trace:
    """Records SSA operations; values are indices into self.ops."""

    def __init__(self):
        self.ops = []            # (opname, a, b, imm)  a/b are value ids or None
        self.emits = []

    def new(self, op, a=None, b=None, imm=None):
        self.ops.append((op, a, b, imm))
        return Sym(self, len(self.ops) - 1)

    def wire(self, i):
        return self.new("LOADW", imm=i)

    def const_col(self, col):
        return self.new("LOADC", imm=col)

    def pi(self, i):
        return self.new("LOADPI", imm=i)

    def konst(self, k):
        return self.new("LOADK", imm=k % P)

    def multi(self, op, srcs, imm=0):
        """12-in / 12-out super-instruction: returns the 12 result values."""
        srcs = [_val(x) for x in srcs]
        assert op in MULTI and len(srcs) == 12 and all(isinstance(x, Sym) for x in srcs)
        self.ops.append((op, tuple(x.id for x in srcs), 12, imm))
        base = len(self.ops)
        for k in range(12):
            self.ops.append(("RES", base - 1, k, None))
        return [Sym(self, base + k) for k in range(12)]

    def sbox7(self, x):
        x = _val(x)
        return self.new("SBOX7", x.id)

    def emit(self, s):
        s = _val(s)
        s = s if isinstance(s, Sym) else self.konst(s)
        self.ops.append(("EMIT", s.id, None, None))


class Sym:
    __slots__ = ("t", "id")

    def __init__(self, t, i):
        self.t, self.id = t, i

    def _bin(self, o, op, opk, swap=False):
        o = _val(o)
        if isinstance(o, Sym):
            return self.t.new(op, o.id, self.id) if swap else self.t.new(op, self.id, o.id)
        return self.t.new(opk, self.id, None, int(o) % P)

    def __add__(self, o):
        o = _val(o)
        if not isinstance(o, Sym) and int(o) % P == 0:
            return self
        return self._bin(o, "ADD", "ADDK")
    __radd__ = __add__

    def __sub__(self, o):
        o = _val(o)
        if not isinstance(o, Sym) and int(o) % P == 0:
            return self
        return self._bin(o, "SUB", "SUBK")

    def __rsub__(self, o):
        return self.t.new("RSUBK", self.id, None, int(o) % P)

    def __mul__(self, o):
        o = _val(o)
        if not isinstance(o, Sym) and int(o) % P == 1:
            return self
        return self._bin(o, "MUL", "MULK")
    __rmul__ = __mul__


class Ref:
    """A wire / constant / public-input reference that is (re)loaded at every use, so a formula that
    mentions 60 wires does not keep 60 registers live."""
    __slots__ = ("load",)

    def __init__(self, load):
        self.load = load

    def __add__(self, o): return self.load() + _val(o)
    def __radd__(self, o): return self.load() + _val(o)
    def __sub__(self, o): return self.load() - _val(o)
    def __rsub__(self, o): return _val(o) - self.load() if isinstance(_val(o), Sym) else (o - self.load())
    def __mul__(self, o): return self.load() * _val(o)
    def __rmul__(self, o): return self.load() * _val(o)


def _val(o):
    return o.load() if isinstance(o, Ref) else o


# ------------------------------------------------------------------------------------------------ gate formulas
SUPEROPS = True          # set by build_program: RANGE4 / MADK / Poseidon super-instructions vs scalar operations only


def _madk(acc, k, l):
    """acc * k + l"""
    acc, l = _val(acc), _val(l)
    if SUPEROPS and isinstance(acc, Sym) and isinstance(l, Sym):
        return acc.t.new("MADK", acc.id, l.id, int(k) % P)
    return acc * k + l


def _horner(limbs, base):
    acc = limbs[-1]
    for l in reversed(limbs[:-1]):
        acc = _madk(acc, base, l)
    return acc


def _range4(x):
    x = _val(x)
    if SUPEROPS:
        return x.t.new("RANGE4", x.id)
    return x * (x - 1) * (x - 2) * (x - 3)


def gate_noop(t, w, c, pi, params):
    pass


def gate_constant(t, w, c, pi, params):
    for i in range(params["num_consts"]):
        t.emit(c(i) - w(i))


def gate_public_input(t, w, c, pi, params):
    for i in range(4):
        t.emit(w(i) - pi(i))


def gate_arithmetic(t, w, c, pi, params):
    c0, c1 = _val(c(0)), _val(c(1))
    for i in range(params["num_ops"]):
        m0, m1, addend, out = w(4 * i), w(4 * i + 1), w(4 * i + 2), w(4 * i + 3)
        t.emit(out - (m0 * m1 * c0 + addend * c1))


def gate_base_sum(t, w, c, pi, params):
    limbs = [w(1 + i) for i in range(params["num_limbs"])]
    t.emit(_horner(limbs, 2) - w(0))
    for l in limbs:
        l = _val(l)
        t.emit(l * (l - 1))


def gate_u32_arithmetic(t, w, c, pi, params):
    n = params["num_ops"]
    for i in range(n):
        m0, m1, addend = w(6 * i), w(6 * i + 1), w(6 * i + 2)
        lo, hi, inv = w(6 * i + 3), w(6 * i + 4), w(6 * i + 5)
        computed = m0 * m1 + addend
        hi_not_max = inv * (0xFFFFFFFF - hi) - 1
        t.emit(hi_not_max * lo)
        t.emit(hi * (1 << 32) + lo - computed)
        low = high = None
        for j in reversed(range(32)):
            limb = w(6 * n + 32 * i + j)
            t.emit(_range4(limb))
            if j < 16:
                low = limb if low is None else _madk(low, 4, limb)
            else:
                high = limb if high is None else _madk(high, 4, limb)
        t.emit(low - lo)
        t.emit(high - hi)


def gate_u32_subtraction(t, w, c, pi, params):
    n = params["num_ops"]
    for i in range(n):
        x, y, b_in, res, b_out = (w(5 * i + k) for k in range(5))
        t.emit(res - (x - y - b_in + b_out * (1 << 32)))
        comb = None
        for j in reversed(range(16)):
            limb = w(5 * n + 16 * i + j)
            t.emit(_range4(limb))
            comb = limb if comb is None else _madk(comb, 4, limb)
        t.emit(comb - res)
        b_out = _val(b_out)
        t.emit(b_out * (1 - b_out))


def gate_u32_range_check(t, w, c, pi, params):
    k = params["num_input_limbs"]
    for i in range(k):
        aux = [w(k + 16 * i + j) for j in range(16)]
        t.emit(_horner(aux, 4) - w(i))
        for l in aux:
            t.emit(_range4(l))


_FAST = None


def poseidon_fast_tables():
    global _FAST
    if _FAST is None:
        d, e = np.zeros(144, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
        k, v, w = np.zeros(22, dtype=np.uint64), np.zeros(242, dtype=np.uint64), np.zeros(242, dtype=np.uint64)
        check(load().vx_poseidon_fast_tables(ptr(d), ptr(e), ptr(k), ptr(v), ptr(w)), "vx_poseidon_fast_tables")
        rc = np.zeros(360, dtype=np.uint64)
        check(load().vx_poseidon_constants(ptr(rc)), "vx_poseidon_constants")
        _FAST = dict(d=[int(x) for x in d], e=[int(x) for x in e], k=[int(x) for x in k], v=[int(x) for x in v],
                     w=[int(x) for x in w], rc=[int(x) for x in rc])
    return _FAST


def gate_poseidon(t, w, c, pi, params):
    """PoseidonGate constraints; with params["superops"] the layers run as the interpreter's native Poseidon
    super-instructions, otherwise as scalar field operations (the same values either way)."""
    T = poseidon_fast_tables()
    rc = T["rc"]
    sup = params.get("superops", True)
    swap = _val(w(24))
    t.emit(swap * (swap - 1))
    delta = [w(25 + i) for i in range(4)]
    for i in range(4):
        t.emit(swap * (w(i + 4) - w(i)) - delta[i])
    st = [w(i) + delta[i] for i in range(4)] + [w(i + 4) - delta[i] for i in range(4)] + [w(i) for i in range(8, 12)]

    def sbox(x):
        if sup:
            return t.sbox7(x)
        x2 = x * x
        x3 = x2 * x
        x4 = x2 * x2
        return x3 * x4

    def mds(s, next_round):
        """MDS * s + RC[next_round] (next_round = 30: nothing added)"""
        if sup:
            return t.multi("MDS12K", s, next_round)
        out = []
        for r in range(12):
            acc = s[r] * (MDS_CIRC[0] + (MDS_DIAG0 if r == 0 else 0))
            for i in range(1, 12):
                acc = acc + s[(i + r) % 12] * MDS_CIRC[i]
            out.append(acc + rc[12 * next_round + r] if next_round < 30 else acc)
        return out

    st = [_val(st[i]) + rc[i] for i in range(12)]
    for r in range(4):                               # first full rounds (constants of round r already added)
        if r != 0:
            for i in range(12):
                sin = w(29 + 12 * (r - 1) + i)
                t.emit(st[i] - sin)
                st[i] = sin
        st = [sbox(x) for x in st]
        if r < 3:
            st = mds(st, r + 1)
        elif sup:                                    # MDS layer merged with the partial rounds' initial matrix
            st = t.multi("DENSE12", st)
        else:
            new = []
            for j in range(12):
                acc = st[0] * T["d"][12 * j]
                for i in range(1, 12):
                    acc = acc + st[i] * T["d"][12 * j + i]
                new.append(acc + T["e"][j])
            st = new
    for r in range(22):                              # partial rounds, sparse form
        sin = w(65 + r)
        t.emit(st[0] - sin)
        if sup:
            st = t.multi("PARTIAL12", [sbox(sin)] + st[1:], r)
            continue
        x0 = sbox(sin) + T["k"][r]
        d = x0 * 25
        for i in range(1, 12):
            d = d + st[i] * T["v"][11 * r + i - 1]
        st = [d] + [st[i] + x0 * T["w"][11 * r + i - 1] for i in range(1, 12)]
    st = [_val(st[i]) + rc[12 * 26 + i] for i in range(12)]
    for r in range(4):                               # second full rounds
        for i in range(12):
            sin = w(87 + 12 * r + i)
            t.emit(st[i] - sin)
            st[i] = sin
        st = mds([sbox(x) for x in st], 27 + r if r < 3 else 30)
    for i in range(12):
        t.emit(st[i] - w(12 + i))


def gate_u32_add_many(t, w, c, pi, params):
    """add_many_u32.rs:149-190"""
    a, n = params["num_addends"], params["num_ops"]
    for i in range(n):
        b = (a + 3) * i
        computed = _val(w(b + a))
        for j in range(a):
            computed = computed + w(b + j)
        res, oc = w(b + a + 1), w(b + a + 2)
        t.emit(oc * (1 << 32) + res - computed)
        comb_res = comb_carry = None
        for j in reversed(range(19)):
            limb = w((a + 3) * n + 19 * i + j)
            t.emit(_range4(limb))
            if j < 16:
                comb_res = limb if comb_res is None else _madk(comb_res, 4, limb)
            else:
                comb_carry = limb if comb_carry is None else _madk(comb_carry, 4, limb)
        t.emit(comb_res - res)
        t.emit(comb_carry - oc)


def _range_product(x, size):
    """prod_{k < size} (x - k)"""
    if size == 4:
        return _range4(x)
    x = _val(x)
    acc = x
    for k in range(1, size):
        acc = acc * (x - k)
    return acc


def gate_comparison(t, w, c, pi, params):
    """comparison.rs:333-410"""
    n, cb = params["num_chunks"], params["chunk_bits"]
    t.emit(_horner([w(4 + i) for i in range(n)], 1 << cb) - w(0))
    t.emit(_horner([w(4 + n + i) for i in range(n)], 1 << cb) - w(1))
    msd = None
    for i in range(n):
        first, second = w(4 + i), w(4 + n + i)
        t.emit(_range_product(first, 1 << cb))
        t.emit(_range_product(second, 1 << cb))
        diff = second - first
        eq, inter = _val(w(4 + 3 * n + i)), _val(w(4 + 4 * n + i))
        t.emit(diff * w(4 + 2 * n + i) - (1 - eq))
        t.emit(eq * diff)
        t.emit(inter - eq * msd if msd is not None else inter)
        msd = inter + (1 - eq) * diff
    t.emit(w(3) - msd)
    bits = [w(4 + 5 * n + i) for i in range(cb + 1)]
    for b in bits:
        b = _val(b)
        t.emit(b * (1 - b))
    t.emit(w(3) + (1 << cb) - _horner(bits, 2))
    t.emit(w(2) - bits[cb])


def _ext_mul(a, b):
    """(a0 + a1 x)(b0 + b1 x) mod x^2 - 7"""
    a0, a1, b0, b1 = _val(a[0]), _val(a[1]), _val(b[0]), _val(b[1])
    return a0 * b0 + a1 * b1 * 7, a0 * b1 + a1 * b0


def gate_arithmetic_extension(t, w, c, pi, params):
    c0, c1 = _val(c(0)), _val(c(1))
    for i in range(params["num_ops"]):
        b = 8 * i
        pr = _ext_mul((w(b), w(b + 1)), (w(b + 2), w(b + 3)))
        for k in range(2):
            t.emit(w(b + 6 + k) - (pr[k] * c0 + w(b + 4 + k) * c1))


def gate_mul_extension(t, w, c, pi, params):
    c0 = _val(c(0))
    for i in range(params["num_ops"]):
        b = 6 * i
        pr = _ext_mul((w(b), w(b + 1)), (w(b + 2), w(b + 3)))
        for k in range(2):
            t.emit(w(b + 4 + k) - pr[k] * c0)


def gate_reducing(t, w, c, pi, params):
    n, ext = params["num_coeffs"], params["ext"]
    alpha = (_val(w(2)), _val(w(3)))
    acc = (w(4), w(5))
    for i in range(n):
        a = 0 if i == n - 1 else 6 + n * (2 if ext else 1) + 2 * i
        pr = _ext_mul(acc, alpha)
        if ext:
            t.emit(pr[0] + w(6 + 2 * i) - w(a))
            t.emit(pr[1] + w(6 + 2 * i + 1) - w(a + 1))
        else:
            t.emit(pr[0] + w(6 + i) - w(a))
            t.emit(pr[1] - w(a + 1))
        acc = (w(a), w(a + 1))


def gate_exponentiation(t, w, c, pi, params):
    n = params["num_power_bits"]
    base = _val(w(0))
    for i in range(n):
        bit = _val(w(1 + (n - 1 - i)))
        factor = bit * base + (1 - bit)
        if i == 0:
            computed = factor
        else:
            prev = _val(w(2 + n + i - 1))
            computed = prev * prev * factor
        t.emit(computed - w(2 + n + i))
    t.emit(w(1 + n) - w(2 + 2 * n - 1))


def gate_poseidon_mds(t, w, c, pi, params):
    for r in range(12):
        for k in range(2):
            acc = w(2 * r + k) * (MDS_CIRC[0] + (MDS_DIAG0 if r == 0 else 0))
            for i in range(1, 12):
                acc = acc + w(2 * ((i + r) % 12) + k) * MDS_CIRC[i]
            t.emit(w(2 * (12 + r) + k) - acc)


def gate_random_access(t, w, c, pi, params):
    bits, copies, extra = params["bits"], params["num_copies"], params["num_extra_constants"]
    vs = 1 << bits
    routed = (2 + vs) * copies + extra
    for cp in range(copies):
        b0 = (2 + vs) * cp
        bw = [_val(w(routed + cp * bits + i)) for i in range(bits)]
        for b in bw:
            t.emit(b * (b - 1))
        t.emit(_horner(bw, 2) - w(b0))
        items = [w(b0 + 2 + i) for i in range(vs)]
        for b in bw:
            nxt = []
            for k in range(len(items) // 2):
                x = _val(items[2 * k])
                nxt.append(x + b * (items[2 * k + 1] - x))
            items = nxt
        t.emit(items[0] - w(b0 + 1))
    for i in range(extra):
        t.emit(c(i) - w((2 + vs) * copies + i))


def _coset_tables(bits):
    """two_adic_subgroup(bits) in natural order and its barycentric weights 1 / prod_{j != i}(x_i - x_j) = x_i / 2^bits."""
    g = pow(7277203076849721926, 1 << (32 - bits), P)
    dom = [pow(g, i, P) for i in range(1 << bits)]
    inv_n = pow(1 << bits, P - 2, P)
    return dom, [x * inv_n % P for x in dom]


def gate_coset_interpolation(t, w, c, pi, params):
    """plonky2 v0.2.0 gates/coset_interpolation.rs (restated; see DESIGN.md gate coverage): barycentric interpolation over
    shift*<w_16> in the quadratic extension with `num_intermediates` wired (eval, prod) checkpoints."""
    bits, deg = params["subgroup_bits"], params["degree"]
    npts = 1 << bits
    ni = (npts - 2) // (deg - 1)
    dom, wts = _coset_tables(bits)
    p_pt = 1 + 2 * npts
    p_val, p_int = p_pt + 2, p_pt + 4
    p_sh = p_int + 4 * ni
    x = (_val(w(p_sh)), _val(w(p_sh + 1)))
    shift = _val(w(0))
    t.emit(w(p_pt) - x[0] * shift)
    t.emit(w(p_pt + 1) - x[1] * shift)

    def fold(lo, hi, ev, pr):
        for i in range(lo, hi):
            term = (x[0] - dom[i], x[1])
            v = (w(1 + 2 * i) * wts[i], w(2 + 2 * i) * wts[i])
            if pr is None:                                   # first point: eval = 0, prod = 1
                ev, pr = v, term
                continue
            vp, et = _ext_mul(v, pr), _ext_mul(ev, term)
            ev = (et[0] + vp[0], et[1] + vp[1])
            pr = _ext_mul(pr, term)
        return ev, pr
    ev, pr = fold(0, deg, None, None)
    for i in range(ni):
        ie, ip = p_int + 2 * i, p_int + 2 * (ni + i)
        t.emit(w(ie) - ev[0])
        t.emit(w(ie + 1) - ev[1])
        t.emit(w(ip) - pr[0])
        t.emit(w(ip + 1) - pr[1])
        s = 1 + (deg - 1) * (i + 1)
        ev, pr = fold(s, min(s + deg - 1, npts), (_val(w(ie)), _val(w(ie + 1))), (_val(w(ip)), _val(w(ip + 1))))
    t.emit(w(p_val) - ev[0])
    t.emit(w(p_val + 1) - ev[1])


# id prefix -> (formula, parameter parser, degree, num_constants, num_constraints)
def _p(name, key):
    import re
    m = re.search(key + r": (\d+)", name)
    return int(m.group(1))


GATES = {
    "NoopGate": (gate_noop, lambda s: {}, lambda p: (0, 0, 0)),
    "ConstantGate": (gate_constant, lambda s: {"num_consts": _p(s, "num_consts")}, lambda p: (1, p["num_consts"], p["num_consts"])),
    "PublicInputGate": (gate_public_input, lambda s: {}, lambda p: (1, 0, 4)),
    "ArithmeticGate": (gate_arithmetic, lambda s: {"num_ops": _p(s, "num_ops")}, lambda p: (3, 2, p["num_ops"])),
    "BaseSumGate": (gate_base_sum, lambda s: {"num_limbs": _p(s, "num_limbs")}, lambda p: (2, 0, 1 + p["num_limbs"])),
    "U32ArithmeticGate": (gate_u32_arithmetic, lambda s: {"num_ops": _p(s, "num_ops")}, lambda p: (4, 0, 36 * p["num_ops"])),
    "U32SubtractionGate": (gate_u32_subtraction, lambda s: {"num_ops": _p(s, "num_ops")}, lambda p: (4, 0, 19 * p["num_ops"])),
    "U32RangeCheckGate": (gate_u32_range_check, lambda s: {"num_input_limbs": _p(s, "num_input_limbs")},
                          lambda p: (4, 0, 17 * p["num_input_limbs"])),
    "PoseidonGate": (gate_poseidon, lambda s: {}, lambda p: (7, 0, 123)),
    "U32AddManyGate": (gate_u32_add_many, lambda s: {"num_addends": _p(s, "num_addends"), "num_ops": _p(s, "num_ops")},
                       lambda p: (4, 0, 22 * p["num_ops"])),
    "ComparisonGate": (gate_comparison,
                       lambda s: {"num_chunks": _p(s, "num_chunks"), "chunk_bits": -(-_p(s, "num_bits") // _p(s, "num_chunks"))},
                       lambda p: (1 << p["chunk_bits"], 0, 6 + 5 * p["num_chunks"] + p["chunk_bits"])),
    "ArithmeticExtensionGate": (gate_arithmetic_extension, lambda s: {"num_ops": _p(s, "num_ops")}, lambda p: (3, 2, 2 * p["num_ops"])),
    "MulExtensionGate": (gate_mul_extension, lambda s: {"num_ops": _p(s, "num_ops")}, lambda p: (3, 1, 2 * p["num_ops"])),
    "ReducingGate": (gate_reducing, lambda s: {"num_coeffs": _p(s, "num_coeffs"), "ext": False}, lambda p: (2, 0, 2 * p["num_coeffs"])),
    "ReducingExtensionGate": (gate_reducing, lambda s: {"num_coeffs": _p(s, "num_coeffs"), "ext": True},
                              lambda p: (2, 0, 2 * p["num_coeffs"])),
    "ExponentiationGate": (gate_exponentiation, lambda s: {"num_power_bits": _p(s, "num_power_bits")},
                           lambda p: (4, 0, p["num_power_bits"] + 1)),
    "PoseidonMdsGate": (gate_poseidon_mds, lambda s: {}, lambda p: (1, 0, 24)),
    "RandomAccessGate": (gate_random_access,
                         lambda s: {"bits": _p(s, "bits"), "num_copies": _p(s, "num_copies"),
                                    "num_extra_constants": _p(s, "num_extra_constants")},
                         lambda p: (p["bits"] + 1, p["num_extra_constants"], p["num_copies"] * (p["bits"] + 2) + p["num_extra_constants"])),
    "CosetInterpolationGate": (gate_coset_interpolation,
                               lambda s: {"subgroup_bits": _p(s, "subgroup_bits"), "degree": _p(s, "degree")},
                               lambda p: (p["degree"], 0, 2 * (2 + 2 * (((1 << p["subgroup_bits"]) - 2) // (p["degree"] - 1))))),
}


def lookup(gate_id: str):
    name = gate_id.split(" ")[0].split("<")[0]
    for prefix, entry in GATES.items():
        if name == prefix:
            fn, parse, meta = entry
            params = parse(gate_id)
            degree, num_constants, num_constraints = meta(params)
            return fn, params, degree, num_constants, num_constraints
    raise KeyError(f"gate {gate_id!r} has no constraint program (see DESIGN.md: gate coverage)")


# ------------------------------------------------------------------------------------------------ assembler
def _pack_regs(regs):
    """12 register numbers -> two operand words (8 + 4 bytes)"""
    w0 = sum(r << (8 * i) for i, r in enumerate(regs[:8]))
    w1 = sum(r << (8 * i) for i, r in enumerate(regs[8:]))
    return [w0, w1]


def assemble(trace: Trace, filter_id):
    """Linear-scan register allocation of one gate's SSA trace -> (bytecode words, filter register)."""
    ops = trace.ops
    last = {}
    for idx, (op, a, b, imm) in enumerate(ops):
        if op == "RES":
            continue
        operands = a if isinstance(a, tuple) else (a, b)
        for v in operands:
            if v is not None:
                last[v] = idx
    if filter_id is not None:
        last[filter_id] = len(ops)                 # stays live until ENDGATE
    free = list(range(NUM_REGS - 1, -1, -1))
    reg = {}
    words = []

    def enc(op, dst=0, a=0, b=0, imm=0):
        return OP[op] | (dst << 8) | (a << 16) | (b << 24) | ((imm & 0xFFFFFFFF) << 32)

    # Column loads are hoisted in groups: when a LOADW / LOADC is reached, the next LOAD_HOIST - 1 loads of the gate are
    # emitted with it (their registers stay allocated until their original last use).  The device issues loads as
    # asynchronous global -> shared copies and waits once per run, so a group costs one memory latency instead of eight.
    emitted = set()

    def emit_load(i):
        lop, _, _, limm = ops[i]
        rd = free.pop()
        reg[i] = rd
        words.append(enc(lop, rd, 0, 0, limm))
        emitted.add(i)
        if i not in last:
            free.append(rd)

    for idx, (op, a, b, imm) in enumerate(ops):
        if op == "RES":
            continue
        if op in ("LOADW", "LOADC"):
            if idx in emitted:
                continue
            if not free:
                raise RuntimeError("gate program needs more than %d registers" % NUM_REGS)
            emit_load(idx)
            k, j = 1, idx + 1
            while k < LOAD_HOIST and j < len(ops) and j < idx + LOAD_WINDOW and len(free) > LOAD_RESERVE:
                if ops[j][0] in ("LOADW", "LOADC") and j not in emitted:
                    emit_load(j)
                    k += 1
                j += 1
            continue
        if op in MULTI:
            srcs = [reg[v] for v in a]
            for v in set(a):
                if last[v] == idx:
                    free.append(reg[v])
            if len(free) < 12:
                raise RuntimeError("gate program needs more than %d registers" % NUM_REGS)
            dsts = []
            for k in range(12):
                rd = free.pop()
                reg[idx + 1 + k] = rd
                dsts.append(rd)
            words += [enc(op, imm=imm)] + _pack_regs(srcs) + _pack_regs(dsts)
            for k in range(12):
                if (idx + 1 + k) not in last:
                    free.append(reg[idx + 1 + k])
            continue
        ra = reg[a] if a is not None else 0
        rb = reg[b] if b is not None else 0
        for v in {a, b}:                           # operands dying here free their register first:
            if v is not None and last[v] == idx:   # the interpreter reads operands before it writes dst
                free.append(reg[v])
        if op == "EMIT":
            words.append(enc("EMIT", 0, ra))
            continue
        if not free:
            raise RuntimeError("gate program needs more than %d registers" % NUM_REGS)
        rd = free.pop()
        reg[idx] = rd
        if op == "LOADPI":
            words.append(enc(op, rd, 0, 0, imm))
        elif op == "LOADK":
            words += [enc(op, rd), imm]
        elif op in ("ADD", "SUB", "MUL"):
            words.append(enc(op, rd, ra, rb))
        elif op in ("SBOX7", "RANGE4"):
            words.append(enc(op, rd, ra))
        elif op == "MADK":
            words += [enc(op, rd, ra, rb), imm]
        else:                                      # ADDK / MULK / RSUBK / SUBK: immediate in the next word
            words += [enc(op, rd, ra), imm]
        if idx not in last:                        # dead value
            free.append(rd)
    return words, (reg[filter_id] if filter_id is not None else 255)


def build_program(gate_ids: list[str], selector_index: list[int], groups: list[tuple], num_selectors: int,
                  superops: bool = True) -> np.ndarray:
    """One program for all gates (already in plonky2's sorted order).  The gate-local constant i is column
    num_selectors + i of the constants_sigmas batch; the filter follows compute_filter().
    superops=False emits scalar field operations only (the same constraint values; kept for A/B and parity tests)."""
    global SUPEROPS
    SUPEROPS = superops
    words = []
    many = num_selectors > 1
    for gi, gid in enumerate(gate_ids):
        fn, params, degree, ncst, ncons = lookup(gid)
        if ncons == 0:
            continue
        t = Trace()

        def w(i, t=t):
            return Ref(lambda: t.wire(i))

        def c(i, t=t):
            return Ref(lambda: t.const_col(num_selectors + i))

        def pi(i, t=t):
            return Ref(lambda: t.pi(i))
        params = dict(params, superops=superops)
        fn(t, w, c, pi, params)
        # filter = prod_{j in group, j != gi} (j - s) [* (UNUSED - s)]
        lo, hi = groups[selector_index[gi]]
        s = t.const_col(selector_index[gi])
        f = None
        for j in range(lo, hi):
            if j != gi:
                term = j - s
                f = term if f is None else f * term
        if many:
            term = 0xFFFFFFFF - s
            f = term if f is None else f * term
        body, freg = assemble(t, f.id if f is not None else None)
        words += [OP["BEGINGATE"]] + body + [OP["ENDGATE"] | (freg << 16)]
    return np.array(words, dtype=np.uint64)
