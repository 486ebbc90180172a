"""CPU tests of the proof wire format (vectorx_b200/proof_io.py): bincode of plonky2's `ProofWithPublicInputs`,
hex-wrapped the way the reference does at contracts/lib/succinctx/plonky2x/core/src/utils/serde/mod.rs:82-111.
The product serializer is checked against the oracle's independent struct.pack statement of the same layout, against
the size the layout implies, and through deserialize -> verify under the oracle verifier."""
import numpy as np
import pytest

import vectorx_b200 as vx
from oracle import plonk, synth


@pytest.fixture(scope="module")
def proven():
    circ, wires, pis = synth.build(6, seed=11)
    proof = plonk.prove(circ, wires, pis)
    assert plonk.verify(circ, proof)
    return circ, proof


def expected_size(circ, proof):
    """Byte count implied by the layout, from the circuit shape alone."""
    cfg = circ.cfg
    cap = 1 << cfg.cap_height
    n_ext = sum(len(v) for v in proof["openings"].values())
    words = 3 * (1 + 4 * cap)                                   # three caps
    words += 9 + 2 * n_ext                                      # nine opening vectors (two lookup vectors empty)
    arities = plonk.fri_reduction_arity_bits(cfg, circ.d)
    words += 1 + len(arities) * (1 + 4 * cap)                   # commit-phase caps
    bits = circ.d + cfg.rate_bits
    widths = [len(r) for r, _ in proof["queries"][0]["initial"]]
    per_query = 1 + sum(1 + w + 1 + 4 * (bits - cfg.cap_height) for w in widths)
    per_query += 1
    size = bits
    for ab in arities:
        size -= ab
        per_query += 1 + 2 * (1 << ab) + 1 + 4 * max(size - cfg.cap_height, 0)
    words += 1 + cfg.num_query_rounds * per_query
    words += 1 + 2 * len(proof["final_poly"]) + 1               # final poly, pow witness
    words += 1 + len(proof["public_inputs"])
    return 8 * words


def test_bytes_match_oracle_statement_and_size(proven):
    circ, proof = proven
    got = vx.proof_to_bytes(proof)
    assert got == plonk.proof_bytes(proof)
    assert len(got) == expected_size(circ, proof)
    hx = vx.proof_to_hex(proof)
    assert hx.startswith("0x") and bytes.fromhex(hx[2:]) == got


def test_layout_landmarks(proven):
    """First and last words sit where bincode puts them: cap length first, public inputs last."""
    circ, proof = proven
    a = np.frombuffer(vx.proof_to_bytes(proof), dtype="<u8")
    cap = 1 << circ.cfg.cap_height
    assert int(a[0]) == cap
    assert np.array_equal(a[1:1 + 4 * cap], np.asarray(proof["wires_cap"], dtype=np.uint64).reshape(-1))
    assert int(a[1 + 4 * cap]) == cap                            # second cap follows immediately
    k = len(proof["public_inputs"])
    assert int(a[-k - 1]) == k and [int(x) for x in a[-k:]] == [int(x) for x in proof["public_inputs"]]
    assert int(a[-k - 2]) == int(proof["pow_witness"])


def test_round_trip_and_verify(proven):
    circ, proof = proven
    data = vx.proof_to_bytes(proof)
    back = vx.proof_from_hex("0x" + data.hex())
    assert vx.proof_to_bytes(back) == data
    assert all(q["x_index"] is None for q in back["queries"])
    assert back["openings"]["lookup_zs"] == [] and back["openings"]["lookup_zs_next"] == []
    from oracle.field import E2
    p = dict(back)
    p["openings"] = {k: [E2(a, b) for a, b in v] for k, v in back["openings"].items()}
    p["final_poly"] = [E2(a, b) for a, b in back["final_poly"]]
    assert plonk.verify(circ, p)                                  # x_index re-derived from the transcript
    flipped = bytearray(data)
    flipped[8 * 3] ^= 1                                           # one bit of the wires cap
    q = vx.proof_from_bytes(bytes(flipped))
    q["openings"] = p["openings"]; q["final_poly"] = p["final_poly"]
    assert not plonk.verify(circ, q)


def test_malformed_bytes_raise(proven):
    circ, proof = proven
    data = vx.proof_to_bytes(proof)
    with pytest.raises(ValueError):
        vx.proof_from_bytes(data[:-8])
    with pytest.raises(ValueError):
        vx.proof_from_bytes(data + b"\0" * 8)
    with pytest.raises(ValueError):
        vx.proof_from_bytes(data[:-3])
    with pytest.raises(ValueError):
        vx.proof_from_hex(data.hex())
