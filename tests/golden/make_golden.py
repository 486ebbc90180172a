"""Generates tests/golden/commit_golden.npz from the CPU oracle.

Provenance: the oracle is pinned by the reference's only golden vector at this boundary
(test_poseidon, contracts/lib/succinctx/plonky2x/core/src/frontend/hash/poseidon/poseidon256.rs:163-202)
and cross-checked against the independent big-int restatement oracle/pyref.py (tests/test_oracle.py).
The reference implementation itself (Rust, plonky2 v0.2.0) cannot be run in this container, so these
vectors are oracle-derived, not reference-derived: "parity unpinned" beyond Poseidon.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
from oracle import pyref  # noqa: E402

out = {}
# Poseidon
out["perm_zero"] = oracle.poseidon([0] * 12)
out["perm_iota"] = oracle.poseidon(list(range(12)))
st = oracle.random_field((16, 12), seed=101)
out["perm_in"] = st
out["perm_out"] = np.stack([oracle.poseidon(s) for s in st])
inputs, kat_out, got, exp = pyref.kat_poseidon256()
assert got == exp
out["kat_in"] = np.array(inputs, dtype=np.uint64)
out["kat_out"] = np.array(kat_out, dtype=np.uint64)
# hash_no_pad at the lengths the path uses (135 wires, 20 Z/pp, 16 quotient, 32 FRI leaf, 7 short tail)
for ln in (0, 1, 4, 5, 7, 8, 9, 16, 20, 32, 135):
    x = oracle.random_field((ln,), seed=200 + ln)
    out[f"hash_in_{ln}"] = x
    out[f"hash_out_{ln}"] = oracle.hash_no_pad(x)
# tiny commit by hand-checkable size: 2^3 rows x 3 cols, rate 1, cap 1
cols = oracle.random_field((3, 8), seed=301)
r = oracle.commit_from_values(cols, 1, 1)
pc, pl, pd, pcap = pyref.commit_from_values([[int(v) for v in c] for c in cols], 1, 1)
assert r["coeffs"].tolist() == pc and r["leaves"].tolist() == pl and r["digests"].tolist() == pd
for k in ("coeffs", "leaves", "digests", "cap"):
    out[f"tiny_{k}"] = r[k]
out["tiny_cols"] = cols
# standard_recursion_config shape at reduced height: 2^10 x 135, rate 3, cap 4 (cap + sampled rows/paths)
cols = oracle.random_field((135, 1 << 10), seed=302)
r = oracle.commit_from_values(cols, 3, 4)
out["std_seed"] = np.array([302], dtype=np.uint64)
out["std_cap"] = r["cap"]
idx = np.array([0, 1, 2, 4095, 4096, 8191, 5000, 1234], dtype=np.uint64)
out["std_idx"] = idx
out["std_rows"] = r["leaves"][idx.astype(np.int64)]
out["std_paths"] = np.stack([oracle.merkle_prove(r["digests"], 1 << 13, 4, int(i)) for i in idx])
out["std_coeffs_col0"] = r["coeffs"][0]
out["std_digests_head"] = r["digests"][:64]
# narrow oracle (hash_or_noop path): 2^6 x 3, rate 3, cap 4; and cap == tree height
cols = oracle.random_field((3, 64), seed=303)
r = oracle.commit_from_values(cols, 3, 4)
out["narrow_cols"] = cols
out["narrow_cap"] = r["cap"]
out["narrow_digests"] = r["digests"]
cols = oracle.random_field((9, 2), seed=304)
r = oracle.commit_from_values(cols, 1, 2)
out["flat_cols"] = cols
out["flat_cap"] = r["cap"]
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "commit_golden.npz"), **out)
print("wrote", len(out), "arrays")
