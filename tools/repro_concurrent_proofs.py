import sys, threading, numpy as np
sys.path.insert(0, '.')
import vectorx_b200 as vx
from oracle import synth
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 16
circ, wires, pis = synth.build(bits, seed=11)
ctx = vx.Context(0)
pc = vx.CircuitData(circ.d, [g.id() for g in circ.gates], circ.selector_index, circ.groups, circ.constants, circ.sigmas, ctx=ctx)
ref = vx.prove(pc, wires, pis)
res = []
def worker():
    for _ in range(5):
        res.append(vx.prove(pc, wires, pis))
ts = [threading.Thread(target=worker) for _ in range(4)]
[t.start() for t in ts]; [t.join() for t in ts]
def diff(a, b):
    out = []
    for k in ("wires_cap", "zs_pp_cap", "quotient_cap"):
        if not np.array_equal(a[k], b[k]): out.append(k)
    for k in a["openings"]:
        if a["openings"][k] != b["openings"][k]: out.append("open:" + k)
    for i, (x, y) in enumerate(zip(a["fri_caps"], b["fri_caps"])):
        if not np.array_equal(x, y): out.append(f"fri_cap{i}")
    if a["final_poly"] != b["final_poly"]: out.append("final_poly")
    if a["pow_witness"] != b["pow_witness"]: out.append("pow")
    return out
bad = 0
for i, r in enumerate(res):
    d = diff(ref, r)
    if d:
        bad += 1
        print(i, d[:6])
print("bad", bad, "of", len(res))
