"""Consumes tests/golden/rust_pin.npz -- outputs of the REAL plonky2 v0.2.0 prover, written by tools/pin_from_rust.sh on a
machine with cargo (this image has none).  Absent file: the tests skip and every parity claim beyond the Poseidon
known-answer tests stays "unpinned" (DESIGN.md section 5).  Present: the CPU oracle must reproduce it bit for bit (and the
GPU path too, -m gpu), and the domain-separator digest decides the hash_pad block (rate 8 vs width 12)."""
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIN = os.path.join(ROOT, "tests", "golden", "rust_pin.npz")
needs_pin = pytest.mark.skipif(not os.path.exists(PIN), reason="tests/golden/rust_pin.npz absent: run tools/pin_from_rust.sh "
                               "where cargo + plonky2@7445ec9 are available")


def _inputs():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "commit_golden.npz"))
    return oracle.random_field((135, 1 << 10), seed=int(gold["std_seed"][0]))


@needs_pin
def test_oracle_reproduces_the_rust_prover():
    pin = np.load(PIN)
    want = oracle.commit_from_values(_inputs(), 3, 4)
    assert np.array_equal(want["cap"], pin["cap"])
    for k, i in enumerate(pin["probes"]):
        assert np.array_equal(want["leaves"][int(i)], pin["rows"][k])
        assert np.array_equal(oracle.merkle_prove(want["digests"], 1 << 13, 4, int(i)), pin["paths"][k])


@needs_pin
def test_hash_pad_block_matches_the_rust_hasher():
    pin = np.load(PIN)
    if "hash_pad_empty" not in pin:
        pytest.skip("pin written without --hash-pad")
    from oracle import plonk
    got = [int(x) for x in pin["hash_pad_empty"]]
    assert got == plonk.hash_pad([]), ("plonky2's hash_pad(&[]) is not the rate-8 padding this repository defaults to; "
                                       f"width-12 padding gives {plonk.hash_pad([], block=12)}")


@needs_pin
@pytest.mark.gpu
def test_gpu_reproduces_the_rust_prover(ctx):
    import vectorx_b200 as vx
    pin = np.load(PIN)
    b = vx.PolynomialBatch.from_values(_inputs(), 3, False, 4, ctx=ctx)
    idx = [int(i) for i in pin["probes"]]
    assert np.array_equal(b.cap.hashes, pin["cap"])
    assert np.array_equal(b.leaves(idx), pin["rows"]) and np.array_equal(b.prove(idx), pin["paths"])
    b.close()
