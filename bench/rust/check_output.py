"""Compares the Rust harness' output for golden_values.bin with the frozen vectors of this repository
(tests/golden/commit_golden.npz: std_cap, std_rows / std_paths at indices 0, 1, 4096 = N/2, 8191 = N-1).
Equality here pins LDE values, leaf order, digests layout and caps of the oracle and the CUDA path to the real plonky2
prover.  With --write the Rust outputs are also saved (tests/golden/rust_pin.npz) for tests/test_rust_pin.py.

usage: python bench/rust/check_output.py golden_out.bin [--write out.npz] [--hash-pad hash_pad.txt] [--timing run.json]"""
import argparse
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PROBES = (0, 1, 4096, 8191)
DEPTH = 13 - 4


def parse(path):
    out = np.fromfile(path, dtype="<u8")
    cap = out[:64].reshape(16, 4)
    pos, rows, paths = 64, [], []
    for _ in PROBES:
        rows.append(out[pos:pos + 135]); pos += 135
        paths.append(out[pos:pos + 4 * DEPTH].reshape(DEPTH, 4)); pos += 4 * DEPTH
    return cap, np.stack(rows), np.stack(paths)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out_bin")
    ap.add_argument("--write")
    ap.add_argument("--hash-pad")
    ap.add_argument("--timing")
    a = ap.parse_args()
    cap, rows, paths = parse(a.out_bin)
    if a.write:
        extra = {}
        if a.hash_pad:
            extra["hash_pad_empty"] = np.array([int(x) for x in open(a.hash_pad).read().split()], dtype=np.uint64)
        if a.timing:
            extra["rayon_melem_per_s"] = np.array([json.loads(open(a.timing).read().strip().splitlines()[-1])["melem_per_s"]])
        np.savez(a.write, cap=cap, rows=rows, paths=paths, probes=np.array(PROBES), **extra)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "commit_golden.npz"))
    assert np.array_equal(cap, gold["std_cap"]), "cap differs from the frozen oracle / GPU cap"
    idx = [int(i) for i in gold["std_idx"]]
    for k, probe in enumerate(PROBES):
        g = idx.index(probe)
        assert np.array_equal(rows[k], gold["std_rows"][g]), f"leaf {probe} differs"
        assert np.array_equal(paths[k], gold["std_paths"][g]), f"Merkle path of leaf {probe} differs"
    print("plonky2 v0.2.0 output equals the frozen vectors: cap, 4 leaves, 4 Merkle paths")


if __name__ == "__main__":
    main()
