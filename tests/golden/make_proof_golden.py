"""Generates tests/golden/proof_golden.json from the CPU oracle prover (oracle/plonk.py): for seeded synthetic circuits,
the circuit digest, the three caps, the PoW witness and the SHA-256 of the bincode proof bytes.

Provenance: oracle-derived, not reference-derived (the reference's plonky2 v0.2.0 prover is Rust and cannot run in this
container; its tree holds no proof fixtures) -- "parity unpinned" beyond the Poseidon KATs, see DESIGN.md section 5.  The
fixture freezes the oracle AND the GPU path against regressions: tests/test_oracle_plonk.py re-derives it on the CPU and
tests/test_gpu_prover.py requires the GPU proof bytes to hash to the same value.

Run from the repo root:  python tests/golden/make_proof_golden.py
"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import plonk, synth  # noqa: E402

CASES = [
    {"name": "default-mix-2p5", "degree_bits": 5, "seed": 21, "mix": None},
    {"name": "all-19-gates-2p6", "degree_bits": 6, "seed": 22, "mix": "ALL"},
]


def run_case(case):
    mix = synth.ALL_KINDS if case["mix"] == "ALL" else None
    circ, wires, pis = synth.build(case["degree_bits"], seed=case["seed"]) if mix is None else \
        synth.build(case["degree_bits"], seed=case["seed"], mix=mix)
    proof = plonk.prove(circ, wires, pis)
    assert plonk.verify(circ, proof)
    data = plonk.proof_bytes(proof)
    return circ, wires, pis, proof, {
        "name": case["name"], "degree_bits": case["degree_bits"], "seed": case["seed"], "mix": case["mix"],
        "gates": [g.id()[:60] for g in circ.gates],
        "circuit_digest": [int(x) for x in circ.circuit_digest],
        "wires_cap0": [int(x) for x in proof["wires_cap"][0]],
        "zs_pp_cap0": [int(x) for x in proof["zs_pp_cap"][0]],
        "quotient_cap0": [int(x) for x in proof["quotient_cap"][0]],
        "pow_witness": int(proof["pow_witness"]),
        "proof_len": len(data),
        "proof_sha256": hashlib.sha256(data).hexdigest(),
    }


if __name__ == "__main__":
    out = [run_case(c)[4] for c in CASES]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "proof_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path, [o["proof_sha256"][:16] for o in out])
