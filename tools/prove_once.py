"""Runs warm-up + one timed proof of a synthetic 2^k-row circuit (test infrastructure builds the circuit); used under ncu."""
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
