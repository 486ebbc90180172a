"""Compares the Rust harness' output for golden_values.bin with the frozen vectors of this repository
(tests/golden/commit_golden.npz: std_cap, std_rows / std_paths at indices 0, 1, 4096 = N/2, 8191 = N-1).
Equality here pins LDE values, leaf order, digests layout and caps of the CUDA path to the real plonky2 prover.
usage: python bench/rust/check_output.py golden_out.bin"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gold = np.load(os.path.join(ROOT, "tests", "golden", "commit_golden.npz"))
out = np.fromfile(sys.argv[1], dtype="<u8")
cap = out[:64].reshape(16, 4)
assert np.array_equal(cap, gold["std_cap"]), "cap differs from the frozen oracle / GPU cap"
idx = [int(i) for i in gold["std_idx"]]
pos = 64
depth = 13 - 4
for probe in (0, 1, 4096, 8191):
    row = out[pos:pos + 135]; pos += 135
    path = out[pos:pos + 4 * depth].reshape(depth, 4); pos += 4 * depth
    k = idx.index(probe)
    assert np.array_equal(row, gold["std_rows"][k]), f"leaf {probe} differs"
    assert np.array_equal(path, gold["std_paths"][k]), f"Merkle path of leaf {probe} differs"
print("plonky2 v0.2.0 output equals the frozen vectors: cap, 4 leaves, 4 Merkle paths")
