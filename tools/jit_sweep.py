"""usage (GPU box): python tools/jit_sweep.py [bits] [tuning ...] -- wall time of vx_quotient (kernel + coset iNTT + 8 MB
download) with the gate program interpreted and compiled under each tuning word (hex: fence<<16 | threads/32<<8 | minb);
every compiled result must equal the interpreter's bit for bit.  Appends to gpurun_out/jit_sweep.jsonl."""
import ctypes, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorx_b200 as vx
from oracle import synth
from vectorx_b200._lib import check, load, ptr

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 16
tunings = [int(x, 16) for x in sys.argv[2:]] or [0x0404, 0x0802, 0x1001, 0x0405, 0x0406]
ctx = vx.Context(0)
circ, wires, pis = synth.build(bits, seed=11)
pc = vx.CircuitData(circ.d, [g.id() for g in circ.gates], circ.selector_index, circ.groups, circ.constants, circ.sigmas, ctx=ctx)
rng = np.random.default_rng(5)
P = 0xFFFFFFFF00000001
rnd = lambda n: (rng.integers(0, P, size=n, dtype=np.uint64))
betas, gammas, alphas, pi = rnd(2), rnd(2), rnd(2), rnd(4)
zpp = np.zeros((2 * (1 + pc.num_partial_products), pc.n), dtype=np.uint64)
check(load().vx_zs_partial_products(ctx.handle, ctypes.byref(pc.desc), ptr(wires), ptr(pc.sigmas), ptr(betas), ptr(gammas), ptr(zpp)), "zpp")
wb = vx.PolynomialBatch.from_values(wires, pc.rate_bits, False, pc.cap_height, ctx=ctx)
zb = vx.PolynomialBatch.from_values(zpp, pc.rate_bits, False, pc.cap_height, ctx=ctx)
q = np.zeros((2, pc.n << pc.rate_bits), dtype=np.uint64)

def run(reps=6):
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        check(load().vx_quotient(ctx.handle, ctypes.byref(pc.desc), pc.constants_sigmas_commitment.handle, wb.handle,
                                 zb.handle, ptr(pi), ptr(betas), ptr(gammas), ptr(alphas), ptr(q)), "vx_quotient")
        best = min(best, (time.perf_counter() - t) * 1e3)
    return best

os.makedirs("gpurun_out", exist_ok=True)
out = open("gpurun_out/jit_sweep.jsonl", "a")
ms = run()
want = q.copy()
rec = {"bits": bits, "form": "interpreter", "ms": ms}
print(rec); out.write(json.dumps(rec) + "\n")
for t in tunings:
    t0 = time.perf_counter()
    pc.compile_gates(t)
    cs = time.perf_counter() - t0
    ms = run()
    rec = {"bits": bits, "form": "compiled", "tuning": hex(t), "threads": ((t >> 8) & 0xff) * 32 or 128, "min_blocks": t & 0xff, "fence": (t >> 16) & 0xfff or 12,
           "opts": os.environ.get("VX_JIT_OPTS", "default"), "compile_s": round(cs, 1), "ms": ms, "equal": bool(np.array_equal(q, want))}
    print(rec); out.write(json.dumps(rec) + "\n"); out.flush()
    assert rec["equal"]
