"""ctypes binding of libvectorx_b200.so (the C ABI in include/vectorx_b200.h).

There is deliberately no fallback: if the shared library is missing or no B200-class GPU is visible
every entry point raises.  Nothing here imports the test oracle.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# VX_B200_LIB: development override used for A/B builds (csrc/Makefile VARIANT=...); the shipped library otherwise
LIB_PATH = os.environ.get("VX_B200_LIB") or os.path.join(_HERE, "libvectorx_b200.so")

u64p = ctypes.POINTER(ctypes.c_uint64)
u32p = ctypes.POINTER(ctypes.c_uint32)
vp = ctypes.c_void_p
c_u64, c_u32, c_i32, c_i64 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int32, ctypes.c_int64

# name -> (restype, argtypes); must list every symbol declared in include/vectorx_b200.h
SIGNATURES = {
    "vx_ctx_create": (c_i32, [c_i32, ctypes.POINTER(vp)]),
    "vx_device_count": (c_i32, []),
    "vx_device_list": (c_i32, [ctypes.POINTER(c_i32), c_i32]),
    "vx_ctx_destroy": (None, [vp]),
    "vx_last_error": (ctypes.c_char_p, []),
    "vx_device_sync": (c_i32, [vp]),
    "vx_ctx_stream": (vp, [vp]),
    "vx_ctx_launch_count": (c_u64, [vp]),
    "vx_ctx_phase_ms": (c_i32, [vp, ctypes.POINTER(ctypes.c_float)]),
    "vx_dev_alloc": (c_i32, [vp, ctypes.c_size_t, ctypes.POINTER(vp)]),
    "vx_dev_free": (None, [vp, vp]),
    "vx_dev_copy": (c_i32, [vp, vp, vp, ctypes.c_size_t]),
    "vx_host_alloc": (c_i32, [ctypes.c_size_t, ctypes.POINTER(vp)]),
    "vx_host_free": (None, [vp]),
    "vx_host_register": (c_i32, [vp, ctypes.c_size_t]),
    "vx_host_unregister": (None, [vp]),
    "vx_commit_from_coeffs_shard": (c_i32, [vp, vp, c_u32, c_u32, c_u32, c_u32, c_u32, c_u32, ctypes.POINTER(vp)]),
    "vx_batch_shard": (c_i32, [vp, u64p]),
    "vx_shard_group_create": (c_i32, [vp, c_u32, c_u32, c_u32, c_u32, c_u32, c_u32, ctypes.POINTER(vp)]),
    "vx_shard_group_ipc_handle": (c_i32, [vp, vp]),
    "vx_shard_group_connect_ipc": (c_i32, [vp, vp]),
    "vx_shard_group_connect_local": (c_i32, [ctypes.POINTER(vp), c_u32]),
    "vx_shard_commit_from_values": (c_i32, [vp, vp, vp, ctypes.POINTER(vp)]),
    "vx_shard_group_coeffs_device": (vp, [vp]),
    "vx_shard_group_cols_per_rank": (c_u32, [vp]),
    "vx_shard_group_free": (None, [vp]),
    "vx_shard_group_set_timeout": (c_i32, [vp, c_u32]),
    "vx_shard_group_set_layout": (c_i32, [vp, c_u32]),
    "vx_shard_group_column_map": (c_i32, [vp, u32p]),
    "vx_commit_from_values": (c_i32, [vp, vp, c_u32, c_u32, c_u32, c_u32, ctypes.POINTER(vp)]),
    "vx_commit_from_coeffs": (c_i32, [vp, vp, c_u32, c_u32, c_u32, c_u32, ctypes.POINTER(vp)]),
    "vx_commit_from_values_cols": (c_i32, [vp, ctypes.POINTER(vp), c_u32, c_u32, c_u32, c_u32, ctypes.POINTER(vp)]),
    "vx_commit_from_coeffs_cols": (c_i32, [vp, ctypes.POINTER(vp), c_u32, c_u32, c_u32, c_u32, ctypes.POINTER(vp)]),
    "vx_commit_from_values_keep": (c_i32, [vp, vp, c_u32, c_u32, c_u32, c_u32, vp, ctypes.POINTER(vp)]),
    "vx_batch_free": (None, [vp]),
    "vx_batch_shape": (c_i32, [vp, u32p]),
    "vx_batch_cap": (c_i32, [vp, vp]),
    "vx_batch_coeffs": (c_i32, [vp, vp]),
    "vx_batch_leaves": (c_i32, [vp, vp, c_u32, vp]),
    "vx_batch_merkle_paths": (c_i32, [vp, vp, c_u32, vp]),
    "vx_batch_download": (c_i32, [vp, vp, vp]),
    "vx_batch_lde_device": (vp, [vp]),
    "vx_batch_coeffs_device": (vp, [vp]),
    "vx_batch_digests_device": (vp, [vp]),
    "vx_merkle_new": (c_i32, [vp, vp, c_u64, c_u32, c_u32, vp, vp, ctypes.POINTER(vp)]),
    "vx_merkle_new_hasher": (c_i32, [vp, c_u32, vp, c_u64, c_u32, c_u32, vp, vp, ctypes.POINTER(vp)]),
    "vx_commit_from_values_hasher": (c_i32, [vp, c_u32, vp, c_u32, c_u32, c_u32, c_u32, ctypes.POINTER(vp)]),
    "vx_commit_from_coeffs_hasher": (c_i32, [vp, c_u32, vp, c_u32, c_u32, c_u32, c_u32, ctypes.POINTER(vp)]),
    "vx_bn128_permute": (c_i32, [vp, vp, c_u64, vp]),
    "vx_bn128_hash": (c_i32, [vp, vp, c_u64, c_u32, c_i32, vp]),
    "vx_bn128_constants": (c_i32, [vp, vp, vp, vp]),
    "vx_tree_prove": (c_i32, [vp, vp, c_u32, vp]),
    "vx_tree_leaves": (c_i32, [vp, vp, c_u32, vp]),
    "vx_tree_cap": (c_i32, [vp, vp]),
    "vx_tree_free": (None, [vp]),
    "vx_poseidon_permute": (c_i32, [vp, vp, c_u64, vp]),
    "vx_hash_no_pad": (c_i32, [vp, vp, c_u64, c_u32, vp]),
    "vx_challenger_permute": (c_i32, [vp]),
    "vx_poseidon_constants": (c_i32, [vp]),
    "vx_poseidon_fast_tables": (c_i32, [vp, vp, vp, vp, vp]),
    "vx_zs_partial_products": (c_i32, [vp, vp, vp, vp, vp, vp, vp]),
    "vx_quotient": (c_i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "vx_quotient_compile": (c_i32, [vp, vp, c_u32]),
    "vx_quotient_is_compiled": (c_i32, [vp, vp]),
    "vx_quotient_discard": (c_i32, [vp, vp]),
    "vx_quotient_jit_source": (c_i64, [vp, vp, c_u64]),
    "vx_quotient_jit_cubin": (c_i64, [vp, c_u32, vp, c_u64]),
    "vx_batch_eval_ext": (c_i32, [vp, vp, vp]),
    "vx_fri_begin": (c_i32, [vp, vp, c_u32, vp, c_u32, vp, ctypes.POINTER(vp)]),
    "vx_fri_commit_layer": (c_i32, [vp, c_u32, c_u32, vp]),
    "vx_fri_fold": (c_i32, [vp, vp]),
    "vx_fri_final_poly": (c_i32, [vp, vp, u32p]),
    "vx_fri_query": (c_i32, [vp, c_u32, vp, c_u32, vp, vp]),
    "vx_fri_free": (None, [vp]),
    "vx_pow_grind": (c_i32, [vp, vp, c_u32, c_u32, vp]),
    "vx_field_op": (c_i32, [vp, c_u32, vp, vp, vp, c_u64, vp]),
    "vx_ntt": (c_i32, [vp, vp, vp, c_u32, c_u32, c_i32, c_u64]),
}

_lib = None


class VxError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library (no GPU needed for loading; compute calls need one)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VxError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().vx_last_error().decode("utf-8", "replace")
        raise VxError(f"{what} failed with code {rc}: {msg}")


def ptr(a) -> int:
    """Address of a numpy array (host) / torch tensor (host or device) / raw int pointer / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], "need C-contiguous uint64"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):      # torch tensor of dtype int64/uint64
        assert a.is_contiguous() and a.element_size() == 8
        return a.data_ptr()
    raise TypeError(type(a))


class Context:
    """One per process and device: owns the stream, twiddle tables and Poseidon constants."""

    def __init__(self, device: int = 0):
        self._h = vp()
        check(load().vx_ctx_create(device, ctypes.byref(self._h)), "vx_ctx_create")
        self.device = device

    @property
    def handle(self):
        return self._h

    def sync(self):
        check(load().vx_device_sync(self._h), "vx_device_sync")

    @property
    def stream(self) -> int:
        return load().vx_ctx_stream(self._h) or 0

    def phase_ms(self) -> dict:
        out = (ctypes.c_float * 5)()
        check(load().vx_ctx_phase_ms(self._h, out), "vx_ctx_phase_ms")
        return dict(zip(("stage", "intt", "lde", "leaf_hash", "tree_levels"), (float(x) for x in out)))

    @property
    def launch_count(self) -> int:
        return load().vx_ctx_launch_count(self._h)

    def close(self):
        if self._h:
            load().vx_ctx_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count() -> int:
    """sm_100-class devices visible to the library (0 without a GPU)."""
    return int(load().vx_device_count())


def device_list() -> list:
    """CUDA ordinals of the devices a Context can be created on (not necessarily 0..count-1 on a mixed box)."""
    n = device_count()
    if n == 0:
        return []
    buf = (c_i32 * n)()
    k = int(load().vx_device_list(buf, n))
    return [int(buf[i]) for i in range(min(k, n))]


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class _PinnedOwner:
    """Keeps a vx_host_alloc allocation alive for the numpy array that views it."""

    def __init__(self, nbytes: int):
        h = vp()
        check(load().vx_host_alloc(nbytes, ctypes.byref(h)), "vx_host_alloc")
        self.ptr, self.nbytes = h.value, nbytes

    def __del__(self):
        try:
            if self.ptr:
                load().vx_host_free(self.ptr)
                self.ptr = 0
        except Exception:
            pass


def pinned_empty(shape) -> np.ndarray:
    """uint64 numpy array in page-locked host memory (vx_host_alloc): fill the witness matrix in place and every upload
    from it runs at PCIe rate instead of being staged by the driver.  Freed when the last view is collected."""
    shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = 1
    for d in shape:
        n *= d
    owner = _PinnedOwner(max(8 * n, 8))
    buf = (ctypes.c_uint64 * max(n, 1)).from_address(owner.ptr)
    buf._vx_owner = owner                       # numpy keeps `buf` as .base, which keeps the allocation
    return np.frombuffer(buf, dtype=np.uint64, count=n).reshape(shape)


class DeviceArray:
    """A (rows, cols) u64 matrix in device memory owned by the library (vx_dev_alloc): what the prover keeps on the GPU
    between phases.  Accepted wherever the binding takes an array (it exposes .shape and data_ptr())."""

    def __init__(self, ctx: Context, shape, _ptr=None, _owner=None):
        self.ctx, self.shape = ctx, tuple(int(x) for x in shape)
        self._owner = _owner
        if _ptr is None:
            h = vp()
            check(load().vx_dev_alloc(ctx.handle, self.nbytes, ctypes.byref(h)), "vx_dev_alloc")
            self._p = h.value
        else:
            self._p = _ptr

    @property
    def nbytes(self) -> int:
        n = 8
        for d in self.shape:
            n *= d
        return n

    @classmethod
    def from_host(cls, ctx: Context, a: np.ndarray) -> "DeviceArray":
        a = np.ascontiguousarray(a, dtype=np.uint64)
        d = cls(ctx, a.shape)
        check(load().vx_dev_copy(ctx.handle, d._p, a.ctypes.data, a.nbytes), "vx_dev_copy")
        return d

    def to_host(self) -> np.ndarray:
        out = np.zeros(self.shape, dtype=np.uint64)
        check(load().vx_dev_copy(self.ctx.handle, out.ctypes.data, self._p, out.nbytes), "vx_dev_copy")
        return out

    def rows(self, first: int, count: int) -> "DeviceArray":
        """borrowed view of rows [first, first + count)"""
        stride = self.nbytes // self.shape[0]
        return DeviceArray(self.ctx, (count,) + self.shape[1:], _ptr=self._p + first * stride, _owner=self)

    def reshape(self, *shape) -> "DeviceArray":
        v = DeviceArray(self.ctx, shape, _ptr=self._p, _owner=self)
        assert v.nbytes == self.nbytes
        return v

    # duck-typing for ptr()
    def data_ptr(self) -> int:
        return self._p

    def is_contiguous(self) -> bool:
        return True

    def element_size(self) -> int:
        return 8

    def close(self):
        if self._p and self._owner is None:
            load().vx_dev_free(self.ctx.handle, self._p)
        self._p = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
