"""Generates tests/golden/bn128_golden.json from the big-int oracle (oracle/bn128.py), which is itself pinned by the
reference KAT (P2X/backend/wrapper/poseidon_bn128.rs:134-181) and -- whenever /root/reference is mounted -- by a
literal-by-literal comparison of the derived tables with P2X/backend/wrapper/poseidon_bn128_constants.rs.

    python tests/golden/make_bn128_golden.py
"""
import json
import os
import random
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bn128  # noqa: E402

REF = "/root/reference/contracts/lib/succinctx/plonky2x/core/src/backend/wrapper/poseidon_bn128_constants.rs"


def reference_literals():
    """(C, S, M, P) decimal literals of the reference file, or None when the reference is not mounted."""
    if not os.path.exists(REF):
        return None
    text = open(REF).read()
    names = ["load_c_constants", "load_s_constants", "load_m_matrix", "load_p_matrix"]
    cuts = [text.index("fn " + n) for n in names] + [len(text)]
    return tuple([int(x) for x in re.findall(r'"(\d+)"', text[cuts[i]:cuts[i + 1]])] for i in range(4))


def main():
    ref = reference_literals()
    C, S, M, Pm = bn128.optimised_constants()
    flat = (C, S, [x for r in M for x in r], [x for r in Pm for x in r])
    checked = False
    if ref is not None:
        assert ref == flat, "derived tables differ from the reference literals"
        checked = True
    rnd = random.Random(20261017)
    gl = lambda: rnd.randrange(bn128.GL_P)
    out = {
        "tables_sha256": bn128.tables_fingerprint(),
        "tables_checked_against_reference_literals": checked,
        "kat": [[[str(x) for x in i], [str(x) for x in o]] for i, o in bn128.KAT],
        "permute": [], "hash_no_pad": [], "hash_or_noop": [], "two_to_one": [], "tree": {},
    }
    for _ in range(8):
        s = [rnd.randrange(bn128.R) for _ in range(4)]
        out["permute"].append([[str(x) for x in s], [str(x) for x in bn128.permute(s)]])
    for ln in [0, 1, 2, 3, 4, 5, 8, 9, 10, 12, 17, 18, 19, 27, 135]:
        v = [gl() for _ in range(ln)]
        out["hash_no_pad"].append([[str(x) for x in v], str(bn128.hash_no_pad(v))])
        out["hash_or_noop"].append([[str(x) for x in v], str(bn128.hash_or_noop(v))])
    for _ in range(4):
        l, r = rnd.randrange(bn128.R), rnd.randrange(bn128.R)
        out["two_to_one"].append([str(l), str(r), str(bn128.two_to_one(l, r))])
    leaves = [[gl() for _ in range(7)] for _ in range(16)]       # random_data(n, 7), poseidon_bn128.rs:193-195
    dg, cap = bn128.merkle_tree(leaves, 1)
    out["tree"] = {"leaves": [[str(x) for x in l] for l in leaves], "cap_height": 1,
                   "digests": [str(x) for x in dg], "cap": [str(x) for x in cap]}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bn128_golden.json")
    json.dump(out, open(path, "w"), indent=0)
    print("wrote", path, "reference literals checked:", checked)


if __name__ == "__main__":
    main()
