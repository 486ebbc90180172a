"""Oracle-side field helpers (TEST INFRASTRUCTURE): vectorised Goldilocks on numpy uint64 arrays (FA),
scalar base-field (FI) and quadratic-extension (E2) values with the same operator surface, so one
statement of each gate's constraint formula serves the per-point quotient oracle (FA over all LDE
points) and the verifier (E2 at zeta).  Restates plonky2_field v0.2.0 semantics."""
from __future__ import annotations

import numpy as np

P = 0xFFFFFFFF00000001
EPS = 0xFFFFFFFF
W = 7                                   # x^2 = 7
_M32 = np.uint64(0xFFFFFFFF)
_EPS = np.uint64(EPS)
_P = np.uint64(P)
_32 = np.uint64(32)


def _canon(a):
    return np.where(a >= _P, a - _P, a)


def _add(a, b):
    with np.errstate(over="ignore"):
        s = a + b
        s = np.where(s < a, s + _EPS, s)       # both canonical: one fix-up, then canonicalise
    return _canon(s)


def _sub(a, b):
    with np.errstate(over="ignore"):
        d = a - b
        d = np.where(a < b, d - _EPS, d)
    return d                                    # a, b canonical -> d canonical


def _mul(a, b):
    with np.errstate(over="ignore"):
        a0, a1 = a & _M32, a >> _32
        b0, b1 = b & _M32, b >> _32
        p00 = a0 * b0
        mid = a0 * b1 + (p00 >> _32)
        mid2 = a1 * b0 + (mid & _M32)
        hi = a1 * b1 + (mid >> _32) + (mid2 >> _32)
        lo = (mid2 << _32) | (p00 & _M32)
        hh, hl = hi >> _32, hi & _M32
        t = lo - hh
        t = np.where(lo < hh, t - _EPS, t)
        m = hl * _EPS
        r = t + m
        r = np.where(r < t, r + _EPS, r)
    return _canon(r)


def _lift(x, like):
    if isinstance(x, FA):
        return x.v
    return np.uint64(int(x) % P)


class FA:
    """Vector of canonical Goldilocks elements."""
    __slots__ = ("v",)
    __array_priority__ = 100

    def __init__(self, v):
        self.v = np.asarray(v, dtype=np.uint64)

    @staticmethod
    def const(k, n):
        return FA(np.full(n, int(k) % P, dtype=np.uint64))

    def __add__(self, o):
        return FA(_add(self.v, _lift(o, self)))
    __radd__ = __add__

    def __sub__(self, o):
        return FA(_sub(self.v, _lift(o, self)))

    def __rsub__(self, o):
        return FA(_sub(np.broadcast_to(_lift(o, self), self.v.shape), self.v))

    def __mul__(self, o):
        return FA(_mul(self.v, _lift(o, self)))
    __rmul__ = __mul__

    def __neg__(self):
        return FA(_sub(np.zeros_like(self.v), self.v))

    def pow(self, e: int):
        r, b = FA.const(1, self.v.shape[0]), self
        while e:
            if e & 1:
                r = r * b
            b = b * b
            e >>= 1
        return r

    def inv(self):
        return self.pow(P - 2)


class FI:
    """Scalar base-field element."""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = int(v) % P

    def __add__(self, o):
        return FI(self.v + _iv(o))
    __radd__ = __add__

    def __sub__(self, o):
        return FI(self.v - _iv(o))

    def __rsub__(self, o):
        return FI(_iv(o) - self.v)

    def __mul__(self, o):
        return FI(self.v * _iv(o))
    __rmul__ = __mul__

    def __neg__(self):
        return FI(-self.v)

    def __eq__(self, o):
        return self.v == _iv(o)

    def __hash__(self):
        return hash(self.v)

    def inv(self):
        return FI(pow(self.v, P - 2, P))


def _iv(o):
    return o.v if isinstance(o, FI) else int(o) % P


class E2:
    """Element a0 + a1 x of F[x]/(x^2 - 7)."""
    __slots__ = ("a", "b")

    def __init__(self, a, b=0):
        self.a, self.b = int(a) % P, int(b) % P

    @staticmethod
    def lift(o):
        if isinstance(o, E2):
            return o
        if isinstance(o, FI):
            return E2(o.v, 0)
        return E2(int(o), 0)

    def __add__(self, o):
        o = E2.lift(o)
        return E2(self.a + o.a, self.b + o.b)
    __radd__ = __add__

    def __sub__(self, o):
        o = E2.lift(o)
        return E2(self.a - o.a, self.b - o.b)

    def __rsub__(self, o):
        return E2.lift(o) - self

    def __mul__(self, o):
        o = E2.lift(o)
        return E2(self.a * o.a + W * self.b * o.b, self.a * o.b + self.b * o.a)
    __rmul__ = __mul__

    def __neg__(self):
        return E2(-self.a, -self.b)

    def __eq__(self, o):
        o = E2.lift(o)
        return self.a == o.a and self.b == o.b

    def __hash__(self):
        return hash((self.a, self.b))

    def __repr__(self):
        return f"E2({self.a:#x}, {self.b:#x})"

    def inv(self):
        norm = (self.a * self.a - W * self.b * self.b) % P
        ni = pow(norm, P - 2, P)
        return E2(self.a * ni, -self.b * ni)

    def pow(self, e: int):
        r, b = E2(1), self
        while e:
            if e & 1:
                r = r * b
            b = b * b
            e >>= 1
        return r

    def limbs(self):
        return [self.a, self.b]


def fpow(a: int, e: int) -> int:
    return pow(a % P, e, P)


def finv(a: int) -> int:
    return pow(a % P, P - 2, P)


def batch_inverse(a: np.ndarray) -> np.ndarray:
    """Inverse of every element of a uint64 array (all non-zero), vectorised Fermat."""
    return FA(a).inv().v
