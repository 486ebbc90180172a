"""Run-time compiled constraint evaluation (vectorx_b200/csrc/quotient_jit.cu), the part that needs no GPU: the CUDA
translation unit generated from a circuit's gate bytecode has one block per gate and one statement per operation, and
NVRTC compiles it to an sm_100a cubin holding the kernel.  (Bit-exactness of what the kernel computes is a GPU test:
tests/test_gpu_prover.py::test_quotient_polys_compiled.)"""
import ctypes

import numpy as np
import pytest

from oracle import synth
from vectorx_b200 import gates as gate_lib
from vectorx_b200._lib import load
from vectorx_b200.prover import VxCircuitDesc

VX_EUNSUPPORTED = -5


def _desc(mix):
    circ, _, _ = synth.build(5, seed=3, mix=mix)
    gate_ids = [g.id() for g in circ.gates]
    prog = gate_lib.build_program(gate_ids, circ.selector_index, circ.groups, len(circ.groups))
    ngc = max(gate_lib.lookup(g)[4] for g in gate_ids)
    k_is = np.zeros(80, dtype=np.uint64)
    d = VxCircuitDesc(5, 3, 135, 80, circ.constants.shape[0], len(circ.groups), 2, 9, 8, ngc, k_is.ctypes.data,
                      prog.ctypes.data, len(prog))
    return d, prog, k_is, gate_ids


def _source(d):
    lib = load()
    n = lib.vx_quotient_jit_source(ctypes.byref(d), None, 0)
    assert n > 0
    buf = ctypes.create_string_buffer(n + 1)
    assert lib.vx_quotient_jit_source(ctypes.byref(d), buf, n + 1) == n
    return buf.value.decode()


def test_generated_source_mirrors_the_program():
    d, prog, _keep, gate_ids = _desc(("arith", "const", "u32range", "poseidon"))
    src = _source(d)
    ops = {v: k for k, v in gate_lib.OP.items()}
    count, pc = {}, 0
    while pc < len(prog):
        op = ops[int(prog[pc]) & 0xff]
        count[op] = count.get(op, 0) + 1
        pc += 2 if op in ("LOADK", "ADDK", "MULK", "RSUBK", "SUBK", "MADK") else 5 if op in ("MDS12K", "DENSE12", "PARTIAL12") else 1
    assert src.count("GlAcc2 h0, h1;") == count["BEGINGATE"] == count["ENDGATE"]
    assert src.count("ldcol(wj, N,") == count["LOADW"] and src.count("ldcol(cj, N,") == count["LOADC"]
    assert src.count("gl_acc2_mad(h0,") == count["EMIT"]
    assert src.count("gl_pow7_cc(") == count["SBOX7"] and src.count("quot_partial12(st,") == count["PARTIAL12"] == 22
    assert src.count("poseidon_mds_add_freq(st,") == count["MDS12K"] and src.count("poseidon_dense_layer<JB>(st,") == 1
    assert 'extern "C" __global__' in src and "quot_prologue(p, tot)" in src and "quot_epilogue(p, q, tot)" in src
    # alpha-power indices restart after the permutation terms in every gate: 2 + 2 * 10 = 22 with two challenges
    assert "A0(22)" in src and "A0(21)" not in src


def test_every_gate_kind_has_a_translation():
    """All 19 gate programs (every opcode of the bytecode) go through the generator; the statement counts follow the
    program.  (Compiling this one takes ~75 s of NVRTC: it runs on the GPU under VX_TEST_JIT_ALL=1.)"""
    d, prog, _keep, gate_ids = _desc(synth.ALL_KINDS)
    assert len(gate_ids) == 19
    src = _source(d)
    ops = {v: k for k, v in gate_lib.OP.items()}
    count, pc = {}, 0
    while pc < len(prog):
        op = ops[int(prog[pc]) & 0xff]
        count[op] = count.get(op, 0) + 1
        pc += 2 if op in ("LOADK", "ADDK", "MULK", "RSUBK", "SUBK", "MADK") else 5 if op in ("MDS12K", "DENSE12", "PARTIAL12") else 1
    assert src.count("GlAcc2 h0, h1;") == count["BEGINGATE"] == 18              # NoopGate has no constraints
    assert src.count("= gl_mul_cc(r") + src.count("y = gl_mul_cc(a, gl_sub(a, 3))") * 2 == \
        count["MUL"] + count.get("MULK", 0) + 2 * count.get("RANGE4", 0)
    assert src.count("gl_mul_add_cc(") == count["MADK"]
    assert src.count("= gl_add(r") == count["ADD"] + count.get("ADDK", 0)
    assert src.count("= gl_sub(") == count["SUB"] + count.get("SUBK", 0) + count.get("RSUBK", 0)
    assert src.count("p.pi_hash[") == count["LOADPI"] == 4
    assert src.count("VX_FENCE(") >= sum(count.values()) // 16


@pytest.mark.parametrize("mix", [("poseidon", "arith", "const", "u32arith", "u32sub", "u32range", "basesum"), synth.ALL_KINDS],
                         ids=["default-mix", "all-19-gates"])
def test_generated_code_evaluates_like_the_bytecode_on_the_cpu(mix, tmp_path):
    """The generated translation unit, compiled for the HOST against tests/jit_host_shim (big-integer field arithmetic,
    textbook Poseidon layers), gives the same filtered, alpha-reduced constraint sums as the reference interpreter of the
    bytecode (tests/test_oracle_plonk.py::run_program, itself held to the oracle's gate formulas) at random points, for
    both challenges.  No GPU: this pins the generator's operands, immediates, statement order and alpha-power indices."""
    import importlib.util
    import os
    import random
    import re
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("oracle_plonk_tests", os.path.join(here, "test_oracle_plonk.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    d, prog, _keep, gate_ids = _desc(mix)
    src = _source(d)
    src, n = re.subn(r"(?m)^#define VX_FENCE\(k\).*$", "#define VX_FENCE(k) ((void)fence)", src)
    assert n == 1
    src += """
extern "C" void jit_host_set_tables(const u64* rc, const u64* dd, const u64* e, const u64* k, const u64* v, const u64* w) {
    memcpy(T_rc, rc, 360 * 8); memcpy(T_d, dd, 144 * 8); memcpy(T_e, e, 12 * 8); memcpy(T_k, k, 22 * 8);
    memcpy(T_v, v, 242 * 8); memcpy(T_w, w, 242 * 8);
}
extern "C" void jit_host_eval(const u64* wires, const u64* cs, const u64* pi, const u64* apow, unsigned num_terms, u64* out) {
    QuotParams p;
    p.cs = cs; p.wires = wires; p.N = 1; p.apow = apow; p.num_terms = num_terms; p.num_challenges = 2; p.out = out;
    for (int i = 0; i < 4; i++) p.pi_hash[i] = pi[i];
    quotient_jit(p);
}
"""
    cu, so = tmp_path / "gates_host.cpp", tmp_path / "gates_host.so"
    cu.write_text(src)
    subprocess.run(["g++", "-O0", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(here, "jit_host_shim"), str(cu), "-o", str(so)],
                   check=True)
    host = ctypes.CDLL(str(so))
    T = gate_lib.poseidon_fast_tables()
    arr = lambda xs: np.array(xs, dtype=np.uint64)
    tabs = [arr(T[k]) for k in ("rc", "d", "e", "k", "v", "w")]
    host.jit_host_set_tables(*[t.ctypes.data_as(ctypes.c_void_p) for t in tabs])
    rnd = random.Random(11)
    P = gate_lib.P
    nterms = 22 + d.num_gate_constraints
    ncols = d.num_constants
    for trial in range(3):
        alphas = [rnd.randrange(P), rnd.randrange(P)]
        apow = [[pow(a, j, P) for j in range(nterms)] for a in alphas]
        w = [rnd.randrange(P) for _ in range(135)]
        cc = [rnd.randrange(P) for _ in range(ncols)]
        pi = [rnd.randrange(P) for _ in range(4)]
        out = np.zeros(2, dtype=np.uint64)
        aw, ac, ap, aa = arr(w), arr(cc), arr(pi), arr(apow[0] + apow[1])
        host.jit_host_eval(*[x.ctypes.data_as(ctypes.c_void_p) for x in (aw, ac, ap, aa)], ctypes.c_uint(nterms),
                           out.ctypes.data_as(ctypes.c_void_p))
        for k in range(2):
            assert int(out[k]) == ref.run_program(prog, w, cc, pi, apow[k], 22)


def test_malformed_programs_are_rejected():
    d, prog, _keep, _ = _desc(("arith",))
    bad = prog.copy()
    bad[0] = 0xff                                            # unknown opcode
    d.program = bad.ctypes.data
    assert load().vx_quotient_jit_source(ctypes.byref(d), None, 0) < 0
    bad = prog[:-1].copy()                                   # ENDGATE missing
    d.program, d.program_len = bad.ctypes.data, len(bad)
    assert load().vx_quotient_jit_source(ctypes.byref(d), None, 0) < 0


def test_nvrtc_compiles_the_kernel_for_sm_100a():
    d, _prog, _keep, _ = _desc(("arith", "const", "u32range"))
    lib = load()
    cap = 16 << 20
    out = ctypes.create_string_buffer(cap)
    n = lib.vx_quotient_jit_cubin(ctypes.byref(d), 0, out, cap)
    if n == VX_EUNSUPPORTED:
        pytest.skip("NVRTC (libnvrtc.so.12) is not installed here")
    assert n > 0, lib.vx_last_error().decode()
    cubin = out.raw[:n]
    assert cubin[:4] == b"\x7fELF" and b"quotient_jit" in cubin
