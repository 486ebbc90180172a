"""Synthetic circuits + satisfying witnesses for the prover tests (TEST INFRASTRUCTURE).

The real VectorX circuits need the Rust witness generator and Avail headers; these
stand-ins use the same config (135 wires / 80 routed, standard_recursion_config) and a gate mix drawn
from the registered gate list (P2X/backend/circuit/serialization/gates.rs:85-107)."""
from __future__ import annotations

import random

import numpy as np

from . import P, hash_no_pad
from .gates import (ArithmeticExtensionGate, ArithmeticGate, BaseSumGate, ComparisonGate, ConstantGate,
                    CosetInterpolationGate,
                    ExponentiationGate, MulExtensionGate, NoopGate, PoseidonGate, PoseidonMdsGate, PublicInputGate,
                    RandomAccessGate, ReducingExtensionGate, ReducingGate, U32AddManyGate, U32ArithmeticGate,
                    U32RangeCheckGate, U32SubtractionGate, NUM_WIRES)

# every gate kind with a constraint program (19 of the 23 registered types)
ALL_KINDS = ("poseidon", "arith", "const", "u32arith", "u32sub", "u32range", "basesum", "u32addmany", "comparison",
             "arithext", "mulext", "reducing", "reducingext", "exp", "poseidonmds", "randomaccess", "cosetinterp")
from .plonk import Circuit, Config


def build(degree_bits: int, seed: int = 1, mix=("poseidon", "arith", "const", "u32arith", "u32sub", "u32range", "basesum"),
          config: Config | None = None):
    """Returns (circuit, wires (135 x n uint64), public_inputs)."""
    rnd = random.Random(seed)
    n = 1 << degree_bits
    gates = [NoopGate(), PublicInputGate(), ConstantGate(2), ArithmeticGate(20), PoseidonGate(),
             U32ArithmeticGate(3), U32SubtractionGate(6), U32RangeCheckGate(7), BaseSumGate(63),
             U32AddManyGate(3), ComparisonGate(32, 16), ArithmeticExtensionGate(), MulExtensionGate(), ReducingGate(),
             ReducingExtensionGate(), ExponentiationGate(), PoseidonMdsGate(), RandomAccessGate(), CosetInterpolationGate()]
    G = {"noop": 0, "pi": 1, "const": 2, "arith": 3, "poseidon": 4, "u32arith": 5, "u32sub": 6, "u32range": 7, "basesum": 8,
         "u32addmany": 9, "comparison": 10, "arithext": 11, "mulext": 12, "reducing": 13, "reducingext": 14, "exp": 15,
         "poseidonmds": 16, "randomaccess": 17, "cosetinterp": 18}
    used = sorted({G[m] for m in mix} | {0, 1})
    gates_used = [gates[i] for i in used]
    gidx = {g: i for i, g in enumerate(used)}
    wires = [[0] * NUM_WIRES for _ in range(n)]
    row_gate = [gidx[0]] * n
    row_consts = [[] for _ in range(n)]
    copies = []
    public_inputs = [rnd.randrange(P) for _ in range(4)]
    pi_hash = [int(x) for x in hash_no_pad(public_inputs)]
    row_gate[0] = gidx[1]
    for i in range(4):
        wires[0][i] = pi_hash[i]
    last_out = None           # (row, col, value) of a routed wire to chain through copy constraints
    body = [m for m in mix]
    for r in range(1, n - 1):
        kind = body[(r - 1) % len(body)] if rnd.random() < 0.9 else "noop"
        if kind == "noop":
            for c in range(NUM_WIRES):            # unconstrained wires may hold anything
                wires[r][c] = rnd.randrange(P) if rnd.random() < 0.05 else 0
            continue
        row_gate[r] = gidx[G[kind]]
        w = wires[r]
        if kind == "poseidon":
            inp = [rnd.randrange(P) for _ in range(12)]
            if last_out is not None:
                inp[0] = last_out[2]
                copies.append(((last_out[0], last_out[1]), (r, 0)))
            out = PoseidonGate.fill_witness(w, inp, swap=rnd.randrange(2))
            last_out = (r, PoseidonGate.W_OUT + 3, out[3])
        elif kind == "arith":
            c0, c1 = rnd.randrange(P), rnd.randrange(P)
            row_consts[r] = [c0, c1]
            for i in range(20):
                m0, m1, add = rnd.randrange(P), rnd.randrange(P), rnd.randrange(P)
                if i == 0 and last_out is not None:
                    m0 = last_out[2]
                    copies.append(((last_out[0], last_out[1]), (r, 0)))
                if i > 0 and rnd.random() < 0.5:          # chain inside the row
                    add = w[4 * (i - 1) + 3]
                    copies.append(((r, 4 * (i - 1) + 3), (r, 4 * i + 2)))
                w[4 * i:4 * i + 4] = [m0, m1, add, (m0 * m1 * c0 + add * c1) % P]
            last_out = (r, 79, w[79])
        elif kind == "const":
            c0, c1 = rnd.randrange(P), rnd.randrange(1 << 20)
            row_consts[r] = [c0, c1]
            w[0], w[1] = c0, c1
        elif kind == "u32arith":
            for i in range(3):
                m0, m1, add = rnd.randrange(1 << 32), rnd.randrange(1 << 32), rnd.randrange(1 << 32)
                v = m0 * m1 + add
                lo, hi = v & 0xFFFFFFFF, v >> 32
                inv = pow((0xFFFFFFFF - hi) % P, P - 2, P)
                w[6 * i:6 * i + 6] = [m0, m1, add, lo, hi, inv]
                for j in range(32):
                    w[18 + 32 * i + j] = (v >> (2 * j)) & 3
        elif kind == "u32sub":
            for i in range(6):
                x, y, b = rnd.randrange(1 << 32), rnd.randrange(1 << 32), rnd.randrange(2)
                res = x - y - b
                bout = 1 if res < 0 else 0
                res += bout << 32
                w[5 * i:5 * i + 5] = [x, y, b, res, bout]
                for j in range(16):
                    w[30 + 16 * i + j] = (res >> (2 * j)) & 3
        elif kind == "u32range":
            for i in range(7):
                v = rnd.randrange(1 << 32)
                w[i] = v
                for j in range(16):
                    w[7 + 16 * i + j] = (v >> (2 * j)) & 3
        elif kind == "basesum":
            v = rnd.randrange(1 << 63)
            w[0] = v
            for j in range(63):
                w[1 + j] = (v >> j) & 1
        elif kind in ("u32addmany", "comparison", "reducing", "reducingext", "exp", "poseidonmds", "cosetinterp"):
            gates[G[kind]].fill_witness(w, rnd)
        elif kind == "arithext":
            c0, c1 = rnd.randrange(P), rnd.randrange(P)
            row_consts[r] = [c0, c1]
            gates[G[kind]].fill_witness(w, rnd, c0, c1)
        elif kind == "mulext":
            c0 = rnd.randrange(P)
            row_consts[r] = [c0]
            gates[G[kind]].fill_witness(w, rnd, c0)
        elif kind == "randomaccess":
            cs = [rnd.randrange(P), rnd.randrange(P)]
            row_consts[r] = cs
            gates[G[kind]].fill_witness(w, rnd, cs)
    warr = np.array(wires, dtype=np.uint64).T.copy()          # (135, n)
    circ = Circuit(degree_bits, gates_used, row_gate, row_consts, copies, config=config, num_public_inputs=4)
    return circ, warr, public_inputs
