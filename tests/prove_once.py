"""Profiling helper (lives under tests/ because the synthetic circuit comes from the test oracle): warm-up + one timed
proof of a 2^k-row circuit, bracketed by cudaProfilerStart/Stop.  Used under ncu:
  VX_PROVE_BITS=16 ncu --set full -k regex:quotient_kernel -s 1 -c 1 -o out python tests/prove_once.py"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorx_b200 as vx
from oracle import synth
bits = int(os.environ.get("VX_PROVE_BITS", "14"))
circ, wires, pis = synth.build(bits, seed=11)
ctx = vx.default_context(0)
pc = vx.CircuitData(circ.d, [g.id() for g in circ.gates], circ.selector_index, circ.groups, circ.constants, circ.sigmas, ctx=ctx)
vx.prove(pc, wires, pis)
import ctypes
ctypes.CDLL("libcudart.so").cudaProfilerStart()
tr = {"intermediates": False}
vx.prove(pc, wires, pis, trace=tr)
ctypes.CDLL("libcudart.so").cudaProfilerStop()
print(json.dumps(tr["phase_ms"]))
