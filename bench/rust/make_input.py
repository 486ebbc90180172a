"""Writes the value matrices the Rust harness reads (run from the repo root):
  golden  : the 2^10 x 135 case whose cap / rows / paths are frozen in tests/golden/commit_golden.npz (seed 302)
  config1 : BASELINE.json config 1 (2^16 x 135, rate 3, cap 4), the bench.py input set 0
"""
import os
import struct

import numpy as np

P = 0xFFFFFFFF00000001


def random_field(shape, seed):
    """The generator the golden fixtures were made with (tests/golden/make_golden.py): numpy PCG64 with rejection."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2**64, size=shape, dtype=np.uint64)
    bad = a >= np.uint64(P)
    while bad.any():
        a[bad] = rng.integers(0, 2**64, size=int(bad.sum()), dtype=np.uint64)
        bad = a >= np.uint64(P)
    return a


def write(path, cols, rate_bits, cap_height):
    c, n = cols.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<4I", c, n.bit_length() - 1, rate_bits, cap_height))
        f.write(np.ascontiguousarray(cols, dtype="<u8").tobytes())
    print(path, cols.shape)


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    write(os.path.join(here, "golden_values.bin"), random_field((135, 1 << 10), seed=302), 3, 4)
    rng = np.random.default_rng(0x5EED0001)
    write(os.path.join(here, "config1_values.bin"),
          rng.integers(0, 0xFFFFFFFF00000001, size=(135, 1 << 16), dtype=np.uint64), 3, 4)
