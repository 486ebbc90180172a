"""ctypes loader for the CPU oracle (oracle/oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; nothing under vectorx_b200/ does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

P = 0xFFFFFFFF00000001
GENERATOR = 14293326489335486720

u64p = ctypes.POINTER(ctypes.c_uint64)


def build(native: bool = False, force: bool = False) -> str:
    """Compile liboracle.so (portable x86-64-v3) or liboracle_native.so (-march=native, for timing)."""
    name = "liboracle_native.so" if native else "liboracle.so"
    out = os.path.join(_HERE, name)
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "prover.c") if os.path.exists(os.path.join(_HERE, f))]
    deps = srcs + [os.path.join(_HERE, "oracle.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    march = "native" if native else "x86-64-v3"
    cmd = ["/usr/bin/gcc", "-O3", f"-march={march}", "-fopenmp", "-fPIC", "-std=gnu11", "-shared",
           "-o", out] + srcs + ["-lm"]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return out


def _ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def lib(native: bool = False):
    global _LIB
    if _LIB is not None and not native:
        return _LIB
    path = build(native=native)
    try:
        L = ctypes.CDLL(path)
    except OSError:
        path = build(native=native, force=True)
        L = ctypes.CDLL(path)
    c = ctypes
    sig = {
        "vxo_add": (c.c_uint64, [c.c_uint64, c.c_uint64]),
        "vxo_sub": (c.c_uint64, [c.c_uint64, c.c_uint64]),
        "vxo_mul": (c.c_uint64, [c.c_uint64, c.c_uint64]),
        "vxo_pow": (c.c_uint64, [c.c_uint64, c.c_uint64]),
        "vxo_inv": (c.c_uint64, [c.c_uint64]),
        "vxo_root_of_unity": (c.c_uint64, [c.c_uint32]),
        "vxo_ext_mul": (None, [u64p, u64p, u64p]),
        "vxo_ext_inv": (None, [u64p, u64p]),
        "vxo_poseidon_constants": (None, [u64p]),
        "vxo_poseidon": (None, [u64p]),
        "vxo_poseidon_naive": (None, [u64p]),
        "vxo_hash_no_pad": (None, [u64p, c.c_size_t, u64p]),
        "vxo_two_to_one": (None, [u64p, u64p, u64p]),
        "vxo_hash_or_noop": (None, [u64p, c.c_size_t, u64p]),
        "vxo_merkle_new": (None, [u64p, c.c_uint64, c.c_uint32, c.c_uint32, u64p, u64p]),
        "vxo_merkle_prove": (None, [u64p, c.c_uint64, c.c_uint32, c.c_uint64, u64p]),
        "vxo_merkle_verify": (c.c_int, [u64p, c.c_uint32, c.c_uint64, u64p, c.c_uint32, u64p]),
        "vxo_fft": (None, [u64p, c.c_uint32]),
        "vxo_ifft": (None, [u64p, c.c_uint32]),
        "vxo_coset_fft": (None, [u64p, c.c_uint32, c.c_uint64]),
        "vxo_coset_ifft": (None, [u64p, c.c_uint32, c.c_uint64]),
        "vxo_fft_ext": (None, [u64p, c.c_uint32]),
        "vxo_ifft_ext": (None, [u64p, c.c_uint32]),
        "vxo_coset_fft_ext": (None, [u64p, c.c_uint32, c.c_uint64]),
        "vxo_commit_from_values": (None, [u64p, c.c_uint32, c.c_uint32, c.c_uint32, c.c_uint32,
                                          u64p, u64p, u64p, u64p]),
        "vxo_commit_from_coeffs": (None, [u64p, c.c_uint32, c.c_uint32, c.c_uint32, c.c_uint32,
                                          u64p, u64p, u64p]),
        "vxo_num_threads": (c.c_int, []),
        "vxo_set_num_threads": (None, [c.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if not native:
        _LIB = L
    return L


# ----------------------------------------------------------------------------- numpy-level helpers
def poseidon(state) -> np.ndarray:
    s = np.array(state, dtype=np.uint64).copy()
    assert s.shape == (12,)
    lib().vxo_poseidon(_ptr(s))
    return s


def hash_no_pad(inputs) -> np.ndarray:
    x = np.ascontiguousarray(np.array(inputs, dtype=np.uint64))
    out = np.zeros(4, dtype=np.uint64)
    lib().vxo_hash_no_pad(_ptr(x), x.size, _ptr(out))
    return out


def two_to_one(l, r) -> np.ndarray:
    l = np.ascontiguousarray(np.array(l, dtype=np.uint64)); r = np.ascontiguousarray(np.array(r, dtype=np.uint64))
    out = np.zeros(4, dtype=np.uint64)
    lib().vxo_two_to_one(_ptr(l), _ptr(r), _ptr(out))
    return out


def merkle_new(leaves: np.ndarray, cap_height: int):
    """leaves: (n, w) uint64 row-major -> (digests (2(n-2^cap), 4), cap (2^cap, 4))."""
    leaves = np.ascontiguousarray(leaves, dtype=np.uint64)
    n, w = leaves.shape
    digests = np.zeros((max(2 * (n - (1 << cap_height)), 0), 4), dtype=np.uint64)
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    dptr = _ptr(digests) if digests.size else ctypes.cast(None, u64p)
    lib().vxo_merkle_new(_ptr(leaves), n, w, cap_height, dptr, _ptr(cap))
    return digests, cap


def merkle_prove(digests: np.ndarray, n: int, cap_height: int, leaf_index: int) -> np.ndarray:
    depth = (n.bit_length() - 1) - cap_height
    sib = np.zeros((depth, 4), dtype=np.uint64)
    if depth:
        lib().vxo_merkle_prove(_ptr(np.ascontiguousarray(digests)), n, cap_height, leaf_index, _ptr(sib))
    return sib


def merkle_verify(leaf: np.ndarray, leaf_index: int, siblings: np.ndarray, cap: np.ndarray) -> bool:
    leaf = np.ascontiguousarray(leaf, dtype=np.uint64)
    siblings = np.ascontiguousarray(siblings, dtype=np.uint64).reshape(-1, 4)
    sp = _ptr(siblings) if siblings.size else ctypes.cast(None, u64p)
    return bool(lib().vxo_merkle_verify(_ptr(leaf), leaf.size, leaf_index, sp, siblings.shape[0],
                                        _ptr(np.ascontiguousarray(cap, dtype=np.uint64))))


def fft(a, inverse=False, shift=None) -> np.ndarray:
    a = np.array(a, dtype=np.uint64).copy()
    log_n = a.size.bit_length() - 1
    assert 1 << log_n == a.size
    L = lib()
    if inverse:
        if shift is None:
            L.vxo_ifft(_ptr(a), log_n)
        else:
            L.vxo_coset_ifft(_ptr(a), log_n, shift)
    elif shift is None:
        L.vxo_fft(_ptr(a), log_n)
    else:
        L.vxo_coset_fft(_ptr(a), log_n, shift)
    return a


def commit_from_values(cols: np.ndarray, rate_bits: int, cap_height: int, want_leaves=True, want_digests=True,
                       native: bool = False):
    """cols: (c, n) uint64 column-major batch -> dict(coeffs (c,n), leaves (N,c), digests, cap)."""
    cols = np.ascontiguousarray(cols, dtype=np.uint64)
    c, n = cols.shape
    log_n = n.bit_length() - 1
    N = n << rate_bits
    coeffs = np.zeros((c, n), dtype=np.uint64)
    leaves = np.zeros((N, c), dtype=np.uint64) if want_leaves else None
    digests = np.zeros((max(2 * (N - (1 << cap_height)), 0), 4), dtype=np.uint64) if want_digests else None
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    null = ctypes.cast(None, u64p)
    lib(native).vxo_commit_from_values(_ptr(cols), c, log_n, rate_bits, cap_height, _ptr(coeffs),
                                       _ptr(leaves) if want_leaves else null,
                                       _ptr(digests) if (want_digests and digests.size) else null, _ptr(cap))
    return {"coeffs": coeffs, "leaves": leaves, "digests": digests, "cap": cap}


def commit_from_coeffs(coeffs: np.ndarray, rate_bits: int, cap_height: int):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64)
    c, n = coeffs.shape
    log_n = n.bit_length() - 1
    N = n << rate_bits
    leaves = np.zeros((N, c), dtype=np.uint64)
    digests = np.zeros((max(2 * (N - (1 << cap_height)), 0), 4), dtype=np.uint64)
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    null = ctypes.cast(None, u64p)
    lib().vxo_commit_from_coeffs(_ptr(coeffs), c, log_n, rate_bits, cap_height, _ptr(leaves),
                                 _ptr(digests) if digests.size else null, _ptr(cap))
    return {"coeffs": coeffs, "leaves": leaves, "digests": digests, "cap": cap}


def random_field(shape, seed: int) -> np.ndarray:
    """Uniform canonical Goldilocks elements (SplitMix-free: numpy PCG64 with rejection)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2**64, size=shape, dtype=np.uint64)
    bad = a >= np.uint64(P)
    while bad.any():
        a[bad] = rng.integers(0, 2**64, size=int(bad.sum()), dtype=np.uint64)
        bad = a >= np.uint64(P)
    return a
