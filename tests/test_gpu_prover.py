"""GPU parity tests (-m gpu) for the rest of the proving path: Z / partial products, quotient evaluation,
opening evaluation, FRI fold-and-commit, proof-of-work, and the whole proof -- each bit-for-bit against the
oracle prover (oracle/plonk.py), and the GPU proof must verify under the oracle's independent verifier."""
import ctypes
import os

import numpy as np
import pytest

import oracle
import vectorx_b200 as vx
from oracle import P, plonk, pyref, synth
from oracle.field import E2
from vectorx_b200._lib import check, load, ptr

pytestmark = pytest.mark.gpu


def product_circuit(circ, ctx):
    """Plain data out of the oracle's circuit object -> the product's CircuitData (what the Rust shim serialises)."""
    return vx.CircuitData(circ.d, [g.id() for g in circ.gates], circ.selector_index, circ.groups, circ.constants,
                          circ.sigmas, ctx=ctx)


@pytest.fixture(scope="module")
def case(ctx):
    circ, wires, pis = synth.build(7, seed=3)
    tr = {}
    proof = plonk.prove(circ, wires, pis, trace=tr)
    assert plonk.verify(circ, proof)
    return circ, wires, pis, proof, tr, product_circuit(circ, ctx)


def e2l(e):
    return [e.a, e.b]


def test_circuit_digest_and_constants_sigmas(case):
    circ, wires, pis, proof, tr, pc = case
    assert pc.circuit_digest == circ.circuit_digest
    assert np.array_equal(pc.constants_sigmas_commitment.cap.hashes, circ.cs_commit["cap"])
    lv, dg = pc.constants_sigmas_commitment.download()        # what CircuitBuild::serialize stores (build.rs:187-242)
    assert np.array_equal(lv, circ.cs_commit["leaves"]) and np.array_equal(dg, circ.cs_commit["digests"])


def test_zs_partial_products(case, ctx):
    circ, wires, pis, proof, tr, pc = case
    out = np.zeros_like(tr["zpp"])
    betas, gammas = np.array(tr["betas"], dtype=np.uint64), np.array(tr["gammas"], dtype=np.uint64)
    check(load().vx_zs_partial_products(ctx.handle, ctypes.byref(pc.desc), ptr(wires), ptr(pc.sigmas),
                                        ptr(betas), ptr(gammas), ptr(out)), "zpp")
    assert np.array_equal(out, tr["zpp"])
    assert (out[:2, 0] == 1).all()


@pytest.mark.parametrize("superops", [True, False])
def test_quotient_polys(case, ctx, superops):
    """compute_quotient_polys bit-for-bit, with the gate program in both forms: native Poseidon / RANGE4 / MADK
    super-instructions (default) and scalar field operations only."""
    circ, wires, pis, proof, tr, pc = case
    check(load().vx_quotient_discard(ctx.handle, ctypes.byref(pc.desc)), "discard")     # this test is about the interpreter
    if not superops:
        pc = vx.CircuitData(circ.d, [g.id() for g in circ.gates], circ.selector_index, circ.groups, circ.constants,
                            circ.sigmas, ctx=ctx, superops=False)
        assert len(pc.program) > 2 * len(case[5].program)
    rate, cap = pc.rate_bits, pc.cap_height
    wb = vx.PolynomialBatch.from_values(wires, rate, False, cap, ctx=ctx)
    zb = vx.PolynomialBatch.from_values(tr["zpp"], rate, False, cap, ctx=ctx)
    N = pc.n << rate
    q = np.zeros((2, N), dtype=np.uint64)
    u = lambda xs: np.array([int(x) for x in xs], dtype=np.uint64)
    pi, betas, gammas, alphas = u(tr["pi_hash"]), u(tr["betas"]), u(tr["gammas"]), u(tr["alphas"])
    check(load().vx_quotient(ctx.handle, ctypes.byref(pc.desc), pc.constants_sigmas_commitment.handle, wb.handle,
                             zb.handle, ptr(pi), ptr(betas), ptr(gammas), ptr(alphas), ptr(q)), "vx_quotient")
    for k in range(2):
        assert np.array_equal(q[k], tr["quotient_full"][k])
    assert np.array_equal(q.reshape(16, pc.n), tr["quotient_coeffs"])


def _quotient(ctx, pc, wires, tr):
    rate, cap = pc.rate_bits, pc.cap_height
    wb = vx.PolynomialBatch.from_values(wires, rate, False, cap, ctx=ctx)
    zb = vx.PolynomialBatch.from_values(tr["zpp"], rate, False, cap, ctx=ctx)
    q = np.zeros((2, pc.n << rate), dtype=np.uint64)
    u = lambda xs: np.array([int(x) for x in xs], dtype=np.uint64)
    pi, betas, gammas, alphas = u(tr["pi_hash"]), u(tr["betas"]), u(tr["gammas"]), u(tr["alphas"])
    check(load().vx_quotient(ctx.handle, ctypes.byref(pc.desc), pc.constants_sigmas_commitment.handle, wb.handle,
                             zb.handle, ptr(pi), ptr(betas), ptr(gammas), ptr(alphas), ptr(q)), "vx_quotient")
    wb.close(); zb.close()
    return q


def test_quotient_polys_compiled(case, ctx):
    """The gate program compiled at circuit-load time (vx_quotient_compile: bytecode -> straight-line CUDA -> NVRTC ->
    sm_100a) gives the oracle's quotient polynomials bit for bit, like the interpreter that ran in the test above.  From
    here on every test of this process that proves a circuit with this gate set runs the compiled kernel."""
    circ, wires, pis, proof, tr, pc = case
    check(load().vx_quotient_discard(ctx.handle, ctypes.byref(pc.desc)), "discard")     # an earlier test file may have compiled it
    assert not load().vx_quotient_is_compiled(ctx.handle, ctypes.byref(pc.desc))
    q_interpreted = _quotient(ctx, pc, wires, tr)
    assert pc.compile_gates() and pc.gates_compiled
    q = _quotient(ctx, pc, wires, tr)
    assert np.array_equal(q, q_interpreted)
    for k in range(2):
        assert np.array_equal(q[k], tr["quotient_full"][k])
    assert np.array_equal(q.reshape(16, pc.n), tr["quotient_coeffs"])


@pytest.mark.skipif(not os.environ.get("VX_TEST_JIT_ALL"), reason="compiles all 19 gate programs (~75 s of NVRTC); set VX_TEST_JIT_ALL=1")
def test_all_gate_kinds_compiled_proof_is_identical(ctx):
    circ, wires, pis = synth.build(7, seed=7, mix=synth.ALL_KINDS)
    pc = product_circuit(circ, ctx)
    want = vx.proof_to_bytes(vx.prove(pc, wires, pis))
    assert pc.compile_gates()
    assert vx.proof_to_bytes(vx.prove(pc, wires, pis)) == want
    pc.close()


def test_opening_evaluation(case, ctx):
    circ, wires, pis, proof, tr, pc = case
    b = vx.PolynomialBatch.from_values(wires[:9], pc.rate_bits, False, pc.cap_height, ctx=ctx)
    z = tr["zeta"]
    out = np.zeros((9, 2), dtype=np.uint64)
    zz = np.array(e2l(z), dtype=np.uint64)
    check(load().vx_batch_eval_ext(b.handle, ptr(zz), ptr(out)), "eval")
    coeffs = b.polynomials
    for j in range(9):
        want = plonk.eval_base_poly_ext(coeffs[j], z)
        assert [int(out[j, 0]), int(out[j, 1])] == e2l(want)


@pytest.mark.parametrize("pos", [0, 3, 7])
def test_pow_grind_smallest_witness(ctx, pos):
    rng = np.random.default_rng(pos)
    st = oracle.random_field((12,), seed=40 + pos)
    w = ctypes.c_uint64(0)
    check(load().vx_pow_grind(ctx.handle, ptr(st), pos, 12, ctypes.byref(w)), "pow")
    s = [int(x) for x in st]
    cand = 0
    while True:
        t = s[:]; t[pos] = cand
        if int(oracle.poseidon(t)[7]) >> 52 == 0:
            break
        cand += 1
    assert w.value == cand


def normalise(proof):
    """Oracle proofs hold E2 objects, product proofs hold [a, b] lists: bring both to plain nested lists."""
    def ext(e):
        return e2l(e) if isinstance(e, E2) else [int(e[0]), int(e[1])]
    out = {"caps": [np.asarray(proof[k]).tolist() for k in ("wires_cap", "zs_pp_cap", "quotient_cap")],
           "openings": {k: [ext(e) for e in v] for k, v in proof["openings"].items()},
           "fri_caps": [np.asarray(c).tolist() for c in proof["fri_caps"]],
           "final_poly": [ext(e) for e in proof["final_poly"]], "pow_witness": int(proof["pow_witness"]),
           "queries": [{"x_index": int(q["x_index"]),
                        "initial": [(np.asarray(r).tolist(), np.asarray(p).tolist()) for r, p in q["initial"]],
                        "steps": [(np.asarray(r).tolist(), np.asarray(p).tolist()) for r, p in q["steps"]]}
                       for q in proof["queries"]]}
    return out


def to_oracle_proof(proof):
    p = dict(proof)
    p["openings"] = {k: [E2(int(e[0]), int(e[1])) for e in v] for k, v in proof["openings"].items()}
    p["final_poly"] = [E2(int(e[0]), int(e[1])) for e in proof["final_poly"]]
    return p


def test_full_proof_equals_oracle_and_verifies(case):
    circ, wires, pis, proof, tr, pc = case
    gtr = {}
    gp = vx.prove(pc, wires, pis, trace=gtr)
    assert gtr["betas"] == tr["betas"] and gtr["gammas"] == tr["gammas"] and gtr["alphas"] == tr["alphas"]
    assert np.array_equal(gtr["zpp"], tr["zpp"])
    assert np.array_equal(gtr["quotient_coeffs"], tr["quotient_coeffs"])
    assert gtr["zeta"] == e2l(tr["zeta"])
    a, b = normalise(gp), normalise(proof)
    for key in ("caps", "openings", "fri_caps", "final_poly", "pow_witness"):
        assert a[key] == b[key], key
    assert a["queries"] == b["queries"]
    # proof BYTES (bincode of ProofWithPublicInputs, serde/mod.rs:82-96): product serializer on the GPU proof ==
    # the oracle's own serializer on the oracle proof; and the bytes read back verify
    data = vx.proof_to_bytes(gp)
    assert data == plonk.proof_bytes(proof)
    assert plonk.verify(circ, to_oracle_proof(vx.proof_from_bytes(data)))
    assert plonk.verify(circ, to_oracle_proof(gp))              # the unchanged (oracle) verifier accepts the GPU proof
    bad = to_oracle_proof(gp)
    bad["final_poly"][0] = bad["final_poly"][0] + 1
    assert not plonk.verify(circ, bad)


@pytest.mark.parametrize("degree_bits,mix", [(5, ("arith", "const")), (6, ("poseidon",)), (8, ("u32arith", "u32sub", "u32range", "basesum")),
                                             (9, ("poseidon", "arith", "const", "u32arith", "u32sub", "u32range", "basesum")),
                                             (6, ("u32addmany", "comparison", "exp", "randomaccess")),
                                             (6, ("arithext", "mulext", "reducing", "reducingext", "poseidonmds")),
                                             (7, synth.ALL_KINDS)],
                         ids=["arith", "poseidon", "u32", "mixed-9", "addmany-cmp-exp-ra", "extension-gates", "all-19-gates"])
def test_proofs_of_other_shapes_verify(ctx, degree_bits, mix):
    circ, wires, pis = synth.build(degree_bits, seed=degree_bits, mix=mix)
    pc = product_circuit(circ, ctx)
    gp = vx.prove(pc, wires, pis)
    assert plonk.verify(circ, to_oracle_proof(gp))
    if degree_bits <= 6:
        op = plonk.prove(circ, wires, pis)
        assert normalise(gp) == normalise(op)
        assert vx.proof_to_bytes(gp) == plonk.proof_bytes(op)


def _golden_cases():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proof_golden.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("gold", _golden_cases(), ids=lambda g: g["name"])
def test_proof_bytes_match_committed_golden(ctx, gold):
    """The GPU proof of the seeded circuit hashes to the committed fixture (tests/golden/make_proof_golden.py) without
    running the oracle prover: circuit digest, caps, PoW witness, byte length and SHA-256 of the bincode bytes."""
    import hashlib
    mix = synth.ALL_KINDS if gold["mix"] == "ALL" else None
    circ, wires, pis = synth.build(gold["degree_bits"], seed=gold["seed"]) if mix is None else \
        synth.build(gold["degree_bits"], seed=gold["seed"], mix=mix)
    pc = product_circuit(circ, ctx)
    assert [int(x) for x in pc.circuit_digest] == gold["circuit_digest"]
    gp = vx.prove(pc, wires, pis)
    assert [int(x) for x in gp["wires_cap"][0]] == gold["wires_cap0"]
    assert [int(x) for x in gp["zs_pp_cap"][0]] == gold["zs_pp_cap0"]
    assert [int(x) for x in gp["quotient_cap"][0]] == gold["quotient_cap0"]
    assert int(gp["pow_witness"]) == gold["pow_witness"]
    data = vx.proof_to_bytes(gp)
    assert len(data) == gold["proof_len"]
    assert hashlib.sha256(data).hexdigest() == gold["proof_sha256"]


def test_concurrent_proofs_on_one_context_are_identical(ctx):
    """Four threads prove the same input on ONE context (its four lanes) at the same time, several rounds: every proof must
    be byte-identical to the single-threaded one.  Guards the stream-ordering rules between lanes: the FRI coefficients
    re-allocated by a fold used to be freed on the stream they were allocated on (another lane's, idle) while the fold
    kernel on this call's stream was still reading them -- 2 of 20 concurrent 2^16-row proofs came out with different
    FRI caps (tools/repro_concurrent_proofs.py)."""
    import threading
    circ, wires, pis = synth.build(15, seed=5)      # big enough for the folds' kernels to outlast the host (2^12 hid the bug)
    pc = vx.CircuitData(circ.d, [g.id() for g in circ.gates], circ.selector_index, circ.groups, circ.constants,
                        circ.sigmas, ctx=ctx)
    want = vx.proof_to_bytes(vx.prove(pc, wires, pis))
    got, errs = [], []

    def worker():
        try:
            for _ in range(6):
                got.append(vx.proof_to_bytes(vx.prove(pc, wires, pis)))
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))
    ts = [threading.Thread(target=worker) for _ in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    pc.close()
    assert not errs, errs
    assert len(got) == 24 and all(g == want for g in got), f"{sum(g != want for g in got)} of {len(got)} concurrent proofs differ"
