import time, sys
sys.path.insert(0, '.')
import vectorx_b200 as vx
from oracle import synth
from vectorx_b200.local_prover import CircuitSpec, LocalProver
circ, wires, pis = synth.build(16, seed=11)
spec = CircuitSpec(circ.d, [g.id() for g in circ.gates], circ.selector_index, circ.groups, circ.constants, circ.sigmas)
for w in (1, 2, 3, 4):
    lp = LocalProver(devices=[0], workers_per_device=w)
    lp.batch_prove(spec, [(wires, pis)] * max(2, w))
    t = time.perf_counter(); lp.batch_prove(spec, [(wires, pis)] * 24); dt = time.perf_counter() - t
    print(w, "workers:", round(24 / dt, 2), "proofs/s", flush=True)
    lp.close()
