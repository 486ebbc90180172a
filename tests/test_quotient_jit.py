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
