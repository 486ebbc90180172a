"""CPU tests: the oracle prover/verifier (oracle/plonk.py) is self-consistent, and the product's host logic
(challenger, gate bytecode tracer/assembler) agrees with the oracle -- no GPU needed.

Mirrors the reference's own strategy: every prover test there is prove -> verify
(e.g. contracts/lib/succinctx/plonky2x/core/src/backend/circuit/build.rs:307-328) and every custom gate runs
random wires through two statements of its constraints (.../u32/gates/arithmetic_u32.rs:467-485)."""
import random

import numpy as np
import pytest

import vectorx_b200 as vx
from oracle import P, plonk, pyref, synth
from oracle.field import E2, FI
from vectorx_b200 import gates as vgates

OP = vgates.OP


@pytest.fixture(scope="module")
def proven():
    circ, wires, pis = synth.build(6, seed=3)
    trace = {}
    proof = plonk.prove(circ, wires, pis, trace=trace)
    return circ, wires, pis, proof, trace


def test_oracle_proof_verifies(proven):
    circ, wires, pis, proof, _ = proven
    assert plonk.verify(circ, proof)


def test_oracle_rejects_tampering(proven):
    circ, wires, pis, proof, _ = proven
    p2 = dict(proof); p2["openings"] = dict(proof["openings"])
    p2["openings"]["wires"] = list(proof["openings"]["wires"]); p2["openings"]["wires"][7] = p2["openings"]["wires"][7] + 1
    assert not plonk.verify(circ, p2)
    p3 = dict(proof); p3["pow_witness"] = proof["pow_witness"] + 1
    assert not plonk.verify(circ, p3)
    p4 = dict(proof); p4["final_poly"] = [proof["final_poly"][0] + 1] + list(proof["final_poly"][1:])
    assert not plonk.verify(circ, p4)
    p5 = dict(proof); cap = proof["wires_cap"].copy(); cap[0, 0] ^= np.uint64(1); p5["wires_cap"] = cap
    assert not plonk.verify(circ, p5)
    p6 = dict(proof); p6["public_inputs"] = [proof["public_inputs"][0] + 1] + proof["public_inputs"][1:]
    assert not plonk.verify(circ, p6)


def test_oracle_rejects_bad_witness():
    circ, wires, pis = synth.build(5, seed=9, mix=("arith", "const", "u32sub"))
    w2 = wires.copy()
    rows = [r for r in range(circ.n) if circ.gates[circ.row_gate[r]].name.startswith("U32SubtractionGate")]
    w2[31, rows[0]] ^= np.uint64(1)                    # break one (unrouted) range-check limb
    proof = plonk.prove(circ, w2, pis)
    assert not plonk.verify(circ, proof)


def test_quotient_has_low_degree_and_matches_at_random_point(proven):
    """SURVEY A self-check (3): vanishing(z) == Z_H(z) * sum_i z^(n i) chunk_i(z) at a random extension point."""
    circ, wires, pis, proof, tr = proven
    n = circ.n
    assert plonk.verify(circ, proof)
    for qv in tr["quotient_full"]:
        assert len(qv) == n << circ.cfg.rate_bits
    z = E2(123456789, 987654321)
    for k, qv in enumerate(tr["quotient_full"]):
        acc = E2(0)
        for c in reversed([int(v) for v in qv]):
            acc = acc * z + c
        # recompute the vanishing polynomial at z from the committed coefficient polynomials
        cs = [plonk.eval_base_poly_ext(c, z) for c in circ.cs_commit["coeffs"]]
        nsel_c = circ.constants.shape[0]
        from oracle import commit_from_values
        wc = commit_from_values(wires, circ.cfg.rate_bits, circ.cfg.cap_height)["coeffs"]
        zc = commit_from_values(tr["zpp"], circ.cfg.rate_bits, circ.cfg.cap_height)["coeffs"]
        wv = [plonk.eval_base_poly_ext(c, z) for c in wc]
        zv = [plonk.eval_base_poly_ext(c, z) for c in zc]
        g_n = pyref.primitive_root_of_unity(circ.d)
        zn = [plonk.eval_base_poly_ext(c, z * g_n) for c in zc[:2]]
        npp = (len(zv) - 2) // 2
        res, zh = plonk.eval_vanishing(circ, z, cs[:nsel_c], cs[nsel_c:], wv, zv[:2], zn,
                                       [zv[2 + kk * npp:2 + (kk + 1) * npp] for kk in range(2)], tr["pi_hash"],
                                       tr["betas"], tr["gammas"], tr["alphas"], E2(1))
        assert res[k] == zh * acc
        break


def test_host_challenger_matches_oracle():
    rnd = random.Random(5)
    a, b = plonk.Challenger(), vx.Challenger()
    for step in range(40):
        xs = [rnd.randrange(2**64) for _ in range(rnd.randrange(1, 13))]
        a.observe_elements(xs); b.observe_elements(xs)
        k = rnd.randrange(0, 5)
        assert a.get_n_challenges(k) == b.get_n_challenges(k)
    ea, eb = a.get_ext_challenge(), b.get_extension_challenge()
    assert [ea.a, ea.b] == eb
    assert vx.hash_no_pad_host(list(range(20))) == pyref.hash_n_to_hash_no_pad(list(range(20)))


def run_program(words, wires, consts_cols, pi_hash, apow, num_perm_terms):
    """Reference interpreter of the gate bytecode (mirrors the device loop) over Python ints."""
    R = [0] * vgates.NUM_REGS
    pc, total, h, cidx = 0, 0, 0, 0
    words = [int(w) for w in words]
    while pc < len(words):
        ins = words[pc]; pc += 1
        op, dst, ra, rb, imm = ins & 0xFF, (ins >> 8) & 0xFF, (ins >> 16) & 0xFF, (ins >> 24) & 0xFF, ins >> 32
        if op == OP["LOADW"]: R[dst] = wires[imm]
        elif op == OP["LOADC"]: R[dst] = consts_cols[imm]
        elif op == OP["LOADPI"]: R[dst] = pi_hash[imm]
        elif op == OP["LOADK"]: R[dst] = words[pc]; pc += 1
        elif op == OP["ADD"]: R[dst] = (R[ra] + R[rb]) % P
        elif op == OP["SUB"]: R[dst] = (R[ra] - R[rb]) % P
        elif op == OP["MUL"]: R[dst] = R[ra] * R[rb] % P
        elif op == OP["ADDK"]: R[dst] = (R[ra] + words[pc]) % P; pc += 1
        elif op == OP["MULK"]: R[dst] = R[ra] * words[pc] % P; pc += 1
        elif op == OP["RSUBK"]: R[dst] = (words[pc] - R[ra]) % P; pc += 1
        elif op == OP["SUBK"]: R[dst] = (R[ra] - words[pc]) % P; pc += 1
        elif op == OP["BEGINGATE"]: h, cidx = 0, num_perm_terms
        elif op == OP["EMIT"]: h = (h + R[ra] * apow[cidx]) % P; cidx += 1
        elif op == OP["ENDGATE"]: total = (total + (1 if ra == 255 else R[ra]) * h) % P
        elif op == OP["NOP"]: pass
        elif op == OP["MADK"]: R[dst] = (R[ra] * words[pc] + R[rb]) % P; pc += 1
        elif op == OP["SBOX7"]: R[dst] = pow(R[ra], 7, P)
        elif op == OP["RANGE4"]: R[dst] = R[ra] * (R[ra] - 1) * (R[ra] - 2) * (R[ra] - 3) % P
        elif op in (OP["MDS12K"], OP["DENSE12"], OP["PARTIAL12"]):
            unpack = lambda w0, w1: [(w0 >> (8 * k)) & 0xFF for k in range(8)] + [(w1 >> (8 * k)) & 0xFF for k in range(4)]
            srcs, dsts = unpack(words[pc], words[pc + 1]), unpack(words[pc + 2], words[pc + 3])
            pc += 4
            st = [R[r] for r in srcs]
            T = vgates.poseidon_fast_tables()
            if op == OP["MDS12K"]:
                out = [(sum(st[(i + r) % 12] * vgates.MDS_CIRC[i] for i in range(12)) + (vgates.MDS_DIAG0 * st[0] if r == 0 else 0)
                        + (T["rc"][12 * imm + r] if imm < 30 else 0)) % P for r in range(12)]
            elif op == OP["DENSE12"]:
                out = [(sum(st[i] * T["d"][12 * j + i] for i in range(12)) + T["e"][j]) % P for j in range(12)]
            else:
                x0 = (st[0] + T["k"][imm]) % P
                out = [(25 * x0 + sum(st[i] * T["v"][11 * imm + i - 1] for i in range(1, 12))) % P]
                out += [(st[i] + x0 * T["w"][11 * imm + i - 1]) % P for i in range(1, 12)]
            for r, v in zip(dsts, out):
                R[r] = v
        else: raise AssertionError(op)
    return total


@pytest.mark.parametrize("mix", [None, synth.ALL_KINDS], ids=["default-mix", "all-19-gates"])
def test_gate_programs_match_oracle_formulas(mix):
    """Random wires / constants (not a satisfying witness): the assembled bytecode and the oracle's direct
    formulas give the same filtered, alpha-reduced constraint sum -- for every gate type at once."""
    circ, wires, pis = synth.build(5, seed=4) if mix is None else synth.build(6, seed=4, mix=mix)
    ids = [g.id() for g in circ.gates]
    prog = vgates.build_program(ids, circ.selector_index, circ.groups, circ.num_selectors)
    rnd = random.Random(8)
    alpha = rnd.randrange(P)
    nterms = 22 + circ.num_gate_constraints
    apow = [pow(alpha, j, P) for j in range(nterms)]
    for trial in range(3):
        w = [rnd.randrange(P) for _ in range(135)]
        ncols = circ.constants.shape[0]
        cc = [rnd.randrange(P) for _ in range(ncols)]
        for s in range(circ.num_selectors):            # selector columns take gate-index-like values on the LDE too
            cc[s] = rnd.randrange(P)
        pi = [rnd.randrange(P) for _ in range(4)]
        got = run_program(prog, w, cc, pi, apow, 22)
        want = 0
        acc = [0] * circ.num_gate_constraints
        for gi, g in enumerate(circ.gates):
            if g.num_constraints == 0:
                continue
            f = circ.filter(gi, FI(cc[circ.selector_index[gi]]))
            f = 1 if f is None else (f.v if isinstance(f, FI) else f)
            cons = g.eval([FI(x) for x in w], [FI(x) for x in cc[circ.num_selectors:]], pi)
            for ci, cv in enumerate(cons):
                acc[ci] = (acc[ci] + cv.v * f) % P
        want = sum(a * apow[22 + i] for i, a in enumerate(acc)) % P
        assert got == want


def test_poseidon_gate_program_is_satisfied_by_the_permutation():
    from oracle.gates import PoseidonGate
    prog = vgates.build_program(["PoseidonGate"], [0], [(0, 1)], 1)
    rnd = random.Random(1)
    row = [0] * 135
    PoseidonGate.fill_witness(row, [rnd.randrange(P) for _ in range(12)], swap=1)
    apow = [pow(rnd.randrange(P), j, P) for j in range(22 + 123)]
    assert run_program(prog, row, [0], [0] * 4, apow, 22) == 0
    row[70] = (row[70] + 1) % P
    assert run_program(prog, row, [0], [0] * 4, apow, 22) != 0


def test_every_gate_with_a_program_is_satisfied_by_its_honest_witness():
    """The gates whose formulas are restated (in-tree sources for U32AddMany / Comparison, SURVEY Appendix B for the
    upstream ones): an honest witness makes every constraint vanish over the base field, the vectorised field and the
    extension field; random wires do not."""
    from oracle import gates as og
    from oracle.field import FA
    rnd = random.Random(5)
    cases = [(og.U32AddManyGate(3), (), ()), (og.U32AddManyGate(5), (), ()), (og.ComparisonGate(32, 16), (), ()),
             (og.ComparisonGate(10, 5), (), ()), (og.ArithmeticExtensionGate(), (11, 13), (11, 13)),
             (og.MulExtensionGate(), (17,), (17,)), (og.ReducingGate(), (), ()), (og.ReducingExtensionGate(), (), ()),
             (og.ExponentiationGate(), (), ()), (og.PoseidonMdsGate(), (), ()), (og.RandomAccessGate(), ((5, 9),), (5, 9)),
             (og.CosetInterpolationGate(), (), ()), (og.CosetInterpolationGate(3, 8), (), ()),
             (og.CosetInterpolationGate(4, 4), (), ())]
    for gate, fill_args, consts in cases:
        for _ in range(4):
            row = [0] * 135
            gate.fill_witness(row, rnd, *fill_args)
            out = gate.eval([FI(x) for x in row], [FI(x) for x in consts], [FI(0)] * 4)
            assert len(out) == gate.num_constraints, gate.name
            assert all(x.v == 0 for x in out), gate.name
            oa = gate.eval([FA(np.array([x, x], dtype=np.uint64)) for x in row],
                           [FA(np.array([x, x], dtype=np.uint64)) for x in consts], [FA(np.zeros(2, dtype=np.uint64))] * 4)
            assert all((np.asarray(x.v) == 0).all() for x in oa), gate.name
            oe = gate.eval([E2(x, 0) for x in row], [E2(x, 0) for x in consts], [E2(0, 0)] * 4)
            assert all(x.a == 0 and x.b == 0 for x in oe), gate.name
        w = [FI(rnd.randrange(P)) for _ in range(135)]
        out = gate.eval(w, [FI(rnd.randrange(P)) for _ in range(4)], [FI(0)] * 4)
        assert any(x.v != 0 for x in out), gate.name
    g = og.ComparisonGate(32, 16)                      # equal inputs and both orders
    for a, b in ((77, 77), (5, 9), (9, 5), (0, (1 << 32) - 1), ((1 << 32) - 1, 0)):
        row = [0] * 135
        g.fill_witness(row, rnd, a=a, b=b)
        assert row[2] == (1 if a <= b else 0)
        assert all(x.v == 0 for x in g.eval([FI(x) for x in row], [], [FI(0)] * 4))


def test_oracle_proof_with_all_19_gate_types_verifies():
    circ, wires, pis = synth.build(6, seed=11, mix=synth.ALL_KINDS)
    assert len(circ.gates) == 19
    proof = plonk.prove(circ, wires, pis)
    assert plonk.verify(circ, proof)
    rows = [r for r in range(circ.n) if circ.gates[circ.row_gate[r]].name.startswith("ComparisonGate")]
    w2 = wires.copy()
    w2[2, rows[0]] ^= np.uint64(1)                     # flip the comparison result
    assert not plonk.verify(circ, plonk.prove(circ, w2, pis))


def test_coset_interpolation_gate_outputs_the_lagrange_interpolant():
    """The wired evaluation value equals a direct Lagrange interpolation of the 16 wired values at point / shift (an
    independent formula), the shape follows `with_max_degree` (16 points, max degree 8 -> degree 6, 2 intermediates, 47
    wires of which 37 routed, 12 constraints), and the weights are x_i / 16."""
    from oracle import gates as og
    g = og.CosetInterpolationGate()
    assert (g.deg, g.ni, g.num_routed(), g.end(), g.num_constraints) == (6, 2, 37, 47, 12)
    assert g.chunks() == [(0, 6), (6, 11), (11, 16)]
    inv16 = pow(16, P - 2, P)
    assert g.weights == [x * inv16 % P for x in g.domain]
    assert vgates._coset_tables(4) == (g.domain, g.weights)
    rnd = random.Random(3)
    for gate in (g, og.CosetInterpolationGate(3, 8), og.CosetInterpolationGate(4, 4), og.CosetInterpolationGate(2, 8)):
        row = [0] * 135
        x, ev = gate.fill_witness(row, rnd)
        assert gate.interpolate_direct(row, x) == ev
        fn, params, degree, ncon, ncs = vgates.lookup(gate.id())
        assert (degree, ncon, ncs) == (gate.degree, 0, gate.num_constraints)
        assert gate.end() <= 135 and gate.num_routed() <= 80


def test_oracle_proofs_match_committed_golden():
    """tests/golden/proof_golden.json re-derived: the oracle prover is deterministic and frozen (any change to the oracle's
    transcript, gate formulas, FRI or byte layout shows up here before it can silently move the GPU parity target)."""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_proof_golden", os.path.join(here, "golden", "make_proof_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(here, "golden", "proof_golden.json")) as f:
        gold = json.load(f)
    assert [g["name"] for g in gold] == [c["name"] for c in mod.CASES]
    for case, want in zip(mod.CASES, gold):
        assert mod.run_case(case)[4] == want


def test_hash_pad_host_matches_oracle_and_is_not_the_zero_digest():
    """circuit_digest's middle slot is hash_pad(domain_separator = []) (pad10*1), not four zeros: the product-side host
    hash and the oracle agree, and the value is what hash_no_pad gives on the padded block."""
    from oracle import hash_no_pad
    from vectorx_b200.challenger import hash_pad_host
    got = hash_pad_host([])
    assert got == plonk.hash_pad([]) == [int(x) for x in hash_no_pad([1, 0, 0, 0, 0, 0, 0, 1])]
    assert any(got)
    assert hash_pad_host([5, 6, 7]) == plonk.hash_pad([5, 6, 7]) == [int(x) for x in hash_no_pad([5, 6, 7, 1, 0, 0, 0, 1])]
    assert hash_pad_host([], block=12) == [int(x) for x in hash_no_pad([1] + [0] * 10 + [1])]
