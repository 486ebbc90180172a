"""Oracle prover + verifier for synthetic plonky2-style circuits (TEST INFRASTRUCTURE).

Restates plonky2 v0.2.0 `prove_with_partition_witness` (plonk/prover.rs), `compute_quotient_polys` /
`eval_vanishing_poly_base_batch` (plonk/vanishing_poly.rs), `PolynomialBatch::prove_openings`
(fri/oracle.rs), `fri_proof` (fri/prover.rs), `Challenger` (iop/challenger.rs) and the matching
verifier, per SURVEY.md Appendix A.  Upstream is not vendored in /root/reference (Cargo.lock:4847-4850);
the reference reaches this code at contracts/lib/succinctx/plonky2x/core/src/backend/circuit/build.rs:69-75.
PARITY UNPINNED by reference data (no golden proof exists in the reference): held by (i) the proof
verifying under the independent verifier below, (ii) algebraic checks in tests/test_oracle_plonk.py.
"""
from __future__ import annotations

import numpy as np

from . import P, GENERATOR, commit_from_coeffs, commit_from_values, hash_no_pad, lib, _ptr, merkle_new, merkle_prove, \
    merkle_verify, poseidon
from . import pyref


def hash_pad(inputs, block=8):
    """Hasher::hash_pad: pad10*1 up to a multiple of `block`, then hash_no_pad.  In-tree statement of the rule:
    contracts/lib/succinctx/plonky2x/core/src/backend/wrapper/plonky2_config.rs:174-182 (pads to the elements absorbed per
    permutation).  Unpinned against plonky2 v0.2.0 itself (rate 8 vs width 12) -- see DESIGN.md section 5."""
    padded = [int(x) % P for x in inputs] + [1]
    while (len(padded) + 1) % block:
        padded.append(0)
    padded.append(1)
    return [int(x) for x in hash_no_pad(padded)]
from .field import E2, FA, FI, finv, fpow
from .gates import NUM_ROUTED, NUM_WIRES, Gate, NoopGate, PublicInputGate

UNUSED_SELECTOR = 0xFFFFFFFF


class Config:
    """CircuitConfig::standard_recursion_config() (P2X/frontend/builder/mod.rs:69)."""
    num_wires = NUM_WIRES
    num_routed_wires = NUM_ROUTED
    rate_bits = 3
    cap_height = 4
    num_challenges = 2
    quotient_degree_factor = 8
    max_degree = 8
    num_query_rounds = 28
    proof_of_work_bits = 16
    arity_bits = 4
    final_poly_bits = 5

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


def bitrev(i, bits):
    return pyref.bitrev(i, bits)


# ------------------------------------------------------------------------------------------------ Challenger (A.6)
class Challenger:
    def __init__(self):
        self.state = [0] * 12
        self.inp: list[int] = []
        self.out: list[int] = []

    def clone(self):
        c = Challenger()
        c.state, c.inp, c.out = self.state[:], self.inp[:], self.out[:]
        return c

    def _duplex(self):
        for i, x in enumerate(self.inp):
            self.state[i] = x
        self.inp = []
        self.state = [int(v) for v in poseidon(self.state)]
        self.out = self.state[:8]

    def observe_element(self, x):
        self.out = []
        self.inp.append(int(x) % P)
        if len(self.inp) == 8:
            self._duplex()

    def observe_elements(self, xs):
        for x in xs:
            self.observe_element(x)

    def observe_hash(self, h):
        self.observe_elements(h)

    def observe_cap(self, cap):
        for h in cap:
            self.observe_hash(h)

    def observe_ext(self, e: E2):
        self.observe_elements(e.limbs())

    def get_challenge(self) -> int:
        if self.inp or not self.out:
            self._duplex()
        return self.out.pop()

    def get_n_challenges(self, n):
        return [self.get_challenge() for _ in range(n)]

    def get_ext_challenge(self) -> E2:
        a = self.get_challenge()
        b = self.get_challenge()
        return E2(a, b)


# ------------------------------------------------------------------------------------------------ circuit
class Circuit:
    """A synthetic circuit: which gate sits on every row, its constants, and the copy constraints."""

    def __init__(self, degree_bits: int, gates: list[Gate], row_gate: list[int], row_consts: list[list[int]],
                 copies: list[tuple], config: Config | None = None, num_public_inputs: int = 0):
        self.cfg = config or Config()
        self.d = degree_bits
        self.n = 1 << degree_bits
        assert len(row_gate) == self.n
        # gates sorted by (degree, id) as CircuitBuilder::build does
        order = sorted(range(len(gates)), key=lambda i: (gates[i].degree, gates[i].id()))
        self.gates = [gates[i] for i in order]
        remap = {old: new for new, old in enumerate(order)}
        self.row_gate = [remap[g] for g in row_gate]
        self.num_gate_constants = max(g.num_constants for g in self.gates)
        self.num_gate_constraints = max(g.num_constraints for g in self.gates)
        self.num_public_inputs = num_public_inputs
        self._selectors()
        nsel = self.num_selectors
        consts = np.zeros((nsel + self.num_gate_constants, self.n), dtype=np.uint64)
        for r in range(self.n):
            g = self.row_gate[r]
            for s in range(nsel):
                lo, hi = self.groups[s]
                consts[s, r] = g if lo <= g < hi else UNUSED_SELECTOR
            for k, v in enumerate(row_consts[r]):
                consts[nsel + k, r] = v % P
        self.constants = consts
        self.k_is = [fpow(GENERATOR, j) for j in range(self.cfg.num_routed_wires)]
        self._sigmas(copies)
        cs = np.concatenate([self.constants, self.sigmas])
        self.constants_sigmas = cs
        self.cs_commit = commit_from_values(cs, self.cfg.rate_bits, self.cfg.cap_height)
        cap = self.cs_commit["cap"]
        # CircuitBuilder::build: hash_no_pad(cap.flatten() ++ hash_pad(domain_separator = []) ++ [degree_bits])
        self.domain_separator_digest = hash_pad([])
        self.circuit_digest = [int(x) for x in hash_no_pad(
            [int(x) for x in cap.reshape(-1)] + self.domain_separator_digest + [self.d])]

    def _selectors(self):
        md = self.cfg.max_degree
        ng = len(self.gates)
        max_gate_degree = max(g.degree for g in self.gates)
        if max_gate_degree + ng - 1 <= md:            # one selector polynomial is enough
            self.groups = [(0, ng)]
            self.selector_index = [0] * ng
        else:
            groups, start = [], 0
            while start < ng:
                size = 0
                while start + size < ng and size + self.gates[start + size].degree < md:
                    size += 1
                assert size > 0
                groups.append((start, start + size))
                start += size
            self.groups = groups
            self.selector_index = [next(s for s, (lo, hi) in enumerate(groups) if lo <= g < hi) for g in range(ng)]
        self.num_selectors = len(self.groups)

    def _sigmas(self, copies):
        n, R = self.n, self.cfg.num_routed_wires
        parent = list(range(n * R))

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a
        for (r1, c1), (r2, c2) in copies:
            assert c1 < R and c2 < R
            a, b = find(c1 * n + r1), find(c2 * n + r2)
            if a != b:
                parent[a] = b
        classes: dict[int, list[int]] = {}
        for i in range(n * R):
            classes.setdefault(find(i), []).append(i)
        nxt = list(range(n * R))
        for members in classes.values():
            for a, b in zip(members, members[1:] + members[:1]):
                nxt[a] = b
        wn = pyref.primitive_root_of_unity(self.d)
        sub = [1] * n
        for i in range(1, n):
            sub[i] = sub[i - 1] * wn % P
        self.subgroup = sub
        sig = np.zeros((R, n), dtype=np.uint64)
        for c in range(R):
            for r in range(n):
                t = nxt[c * n + r]
                sig[c, r] = self.k_is[t // n] * sub[t % n] % P
        self.sigmas = sig
        self.copy_classes = [m for m in classes.values() if len(m) > 1]

    def filter(self, gate_idx, sel_value):
        """compute_filter: prod_{j in group, j != i} (j - s)  [* (UNUSED - s) with several groups]."""
        lo, hi = self.groups[self.selector_index[gate_idx]]
        f = None
        for j in range(lo, hi):
            if j != gate_idx:
                t = (j - sel_value)
                f = t if f is None else f * t
        if self.num_selectors > 1:
            t = (UNUSED_SELECTOR - sel_value)
            f = t if f is None else f * t
        return f


# ------------------------------------------------------------------------------------------------ vanishing poly (A.9)
def eval_vanishing(circ: Circuit, x, consts, sigmas, wires, zs, zs_next, pps, pi_hash, betas, gammas, alphas, one):
    """Terms and alpha-reduction at point(s) x.  Values are FA (vector) or E2 (scalar). `one` lifts 1."""
    cfg = circ.cfg
    n = circ.n
    nsel = circ.num_selectors
    R = cfg.num_routed_wires
    chunk = cfg.max_degree
    nchunks = (R + chunk - 1) // chunk
    zh = x.pow(n) - 1
    l0 = zh * ((x - 1) * n).inv()
    terms = []
    for k in range(cfg.num_challenges):
        terms.append(l0 * (zs[k] - 1))
    for k in range(cfg.num_challenges):
        acc = [zs[k]] + list(pps[k]) + [zs_next[k]]
        for cidx in range(nchunks):
            num = den = None
            for j in range(cidx * chunk, min(R, (cidx + 1) * chunk)):
                a = wires[j] + x * (betas[k] * circ.k_is[j] % P) + gammas[k]
                b = wires[j] + sigmas[j] * betas[k] + gammas[k]
                num = a if num is None else num * a
                den = b if den is None else den * b
            terms.append(acc[cidx] * num - acc[cidx + 1] * den)
    gate_acc = [None] * circ.num_gate_constraints
    gconsts = consts[nsel:]
    for gi, g in enumerate(circ.gates):
        if g.num_constraints == 0:
            continue
        f = circ.filter(gi, consts[circ.selector_index[gi]])
        cons = g.eval(wires, gconsts, pi_hash)
        for ci, cv in enumerate(cons):
            t = cv if f is None else cv * f
            gate_acc[ci] = t if gate_acc[ci] is None else gate_acc[ci] + t
    zero = one - one
    terms += [zero if t is None else t for t in gate_acc]
    results = []
    for k in range(cfg.num_challenges):
        acc = None
        for t in reversed(terms):               # reduce_with_powers
            acc = t if acc is None else acc * alphas[k] + t
        results.append(acc)
    return results, zh


# ------------------------------------------------------------------------------------------------ prover
def partial_products_and_zs(circ: Circuit, wires: np.ndarray, betas, gammas):
    """A.8: columns [Z_0, Z_1, pp_{0,0..8}, pp_{1,0..8}] over the trace domain."""
    cfg, n = circ.cfg, circ.n
    R, chunk = cfg.num_routed_wires, cfg.max_degree
    nchunks = (R + chunk - 1) // chunk
    sub = FA(np.array(circ.subgroup, dtype=np.uint64))
    zcols, ppcols = [], []
    for k in range(cfg.num_challenges):
        quot = []
        for cidx in range(nchunks):
            num = den = None
            for j in range(cidx * chunk, min(R, (cidx + 1) * chunk)):
                a = FA(wires[j]) + sub * (betas[k] * circ.k_is[j] % P) + gammas[k]
                b = FA(wires[j]) + FA(circ.sigmas[j]) * betas[k] + gammas[k]
                num = a if num is None else num * a
                den = b if den is None else den * b
            quot.append((num * den.inv()).v)
        z = np.zeros(n, dtype=np.uint64)
        pp = np.zeros((nchunks - 1, n), dtype=np.uint64)
        acc = 1
        for r in range(n):                       # sequential running product, as upstream
            z[r] = acc
            for cidx in range(nchunks):
                acc = acc * int(quot[cidx][r]) % P
                if cidx < nchunks - 1:
                    pp[cidx, r] = acc
        assert acc == 1, "copy constraints are not satisfied by the witness"
        zcols.append(z)
        ppcols.append(pp)
    return np.concatenate([np.stack(zcols)] + ppcols)


def eval_base_poly_ext(coeffs: np.ndarray, z: E2) -> E2:
    acc = E2(0)
    for c in reversed([int(v) for v in coeffs]):
        acc = acc * z + c
    return acc


def ext_coset_fft(coeffs: list[E2], log_n: int, shift: int) -> list[E2]:
    buf = np.zeros(2 << log_n, dtype=np.uint64)
    for i, c in enumerate(coeffs):
        buf[2 * i], buf[2 * i + 1] = c.a, c.b
    lib().vxo_coset_fft_ext(_ptr(buf), log_n, shift)
    return [E2(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(1 << log_n)]


def fri_reduction_arity_bits(cfg: Config, degree_bits: int) -> list[int]:
    """ConstantArityBits(4, 5)."""
    out, d = [], degree_bits
    while d > cfg.final_poly_bits and d + cfg.rate_bits - cfg.arity_bits >= cfg.cap_height:
        out.append(cfg.arity_bits)
        d -= cfg.arity_bits
    return out


def pow_check(ch: Challenger, witness: int, bits: int) -> bool:
    c = ch.clone()
    c.observe_element(witness)
    return (c.get_challenge() >> (64 - bits)) == 0


def pow_grind(ch: Challenger, bits: int) -> int:
    """Smallest witness (== the reference under RAYON_NUM_THREADS=1; upstream uses find_any)."""
    st = ch.state[:]
    for i, x in enumerate(ch.inp):
        st[i] = x
    pos = len(ch.inp)
    w = 0
    while True:
        s = st[:]
        s[pos] = w
        if (int(poseidon(s)[7]) >> (64 - bits)) == 0:
            return w
        w += 1


def prove(circ: Circuit, wires: np.ndarray, public_inputs: list[int], trace: dict | None = None):
    """Returns the proof as a dict of plain ints / arrays. `trace` (optional) collects intermediates."""
    cfg = circ.cfg
    n, d, rate = circ.n, circ.d, cfg.rate_bits
    N, bits = n << rate, d + rate
    T = trace if trace is not None else {}
    pi_hash = [int(x) for x in hash_no_pad(public_inputs)]
    ch = Challenger()
    ch.observe_hash(circ.circuit_digest)
    ch.observe_hash(pi_hash)
    wires_c = commit_from_values(wires, rate, cfg.cap_height)
    ch.observe_cap(wires_c["cap"])
    betas = ch.get_n_challenges(cfg.num_challenges)
    gammas = ch.get_n_challenges(cfg.num_challenges)
    zpp = partial_products_and_zs(circ, wires, betas, gammas)
    zpp_c = commit_from_values(zpp, rate, cfg.cap_height)
    ch.observe_cap(zpp_c["cap"])
    alphas = ch.get_n_challenges(cfg.num_challenges)
    T.update(betas=betas, gammas=gammas, alphas=alphas, zpp=zpp, pi_hash=pi_hash)

    # ---- quotient over the LDE coset, natural order (leaves are bit-reversed)
    nat = np.array([bitrev(i, bits) for i in range(N)], dtype=np.int64)     # leaf index of LDE point i
    cs_l, w_l, z_l = circ.cs_commit["leaves"][nat], wires_c["leaves"][nat], zpp_c["leaves"][nat]
    nsel_c = circ.constants.shape[0]
    consts = [FA(cs_l[:, j]) for j in range(nsel_c)]
    sigmas = [FA(cs_l[:, nsel_c + j]) for j in range(cfg.num_routed_wires)]
    wv = [FA(w_l[:, j]) for j in range(cfg.num_wires)]
    nch = cfg.num_challenges
    npp = (zpp.shape[0] - nch) // nch
    zs = [FA(z_l[:, k]) for k in range(nch)]
    nxt = (np.arange(N) + (1 << rate)) % N
    zs_next = [FA(z_l[nxt, k]) for k in range(nch)]
    pps = [[FA(z_l[:, nch + k * npp + i]) for i in range(npp)] for k in range(nch)]
    wN = pyref.primitive_root_of_unity(bits)
    xs = np.zeros(N, dtype=np.uint64)
    acc = GENERATOR
    for i in range(N):
        xs[i] = acc
        acc = acc * wN % P
    x = FA(xs)
    res, zh = eval_vanishing(circ, x, consts, sigmas, wv, zs, zs_next, pps, [FI(v).v for v in pi_hash],
                             betas, gammas, alphas, FA.const(1, N))
    zh_inv = zh.inv()
    qcoeffs = []
    L = lib()
    for k in range(nch):
        qv = (res[k] * zh_inv).v.copy()
        L.vxo_coset_ifft(_ptr(qv), bits, GENERATOR)
        T.setdefault("quotient_full", []).append(qv.copy())
        for c in range(cfg.quotient_degree_factor):
            qcoeffs.append(qv[c * n:(c + 1) * n])
    qcoeffs = np.stack(qcoeffs)
    q_c = commit_from_coeffs(qcoeffs, rate, cfg.cap_height)
    ch.observe_cap(q_c["cap"])
    zeta = ch.get_ext_challenge()
    g_n = pyref.primitive_root_of_unity(d)
    assert zeta.pow(n) != E2(1)
    zeta_next = zeta * g_n

    oracles = [circ.cs_commit, wires_c, zpp_c, q_c]
    def ev(batch, z):
        return [eval_base_poly_ext(c, z) for c in batch["coeffs"]]
    cs_open = ev(circ.cs_commit, zeta)
    openings = {
        "constants": cs_open[:nsel_c], "plonk_sigmas": cs_open[nsel_c:], "wires": ev(wires_c, zeta),
        "plonk_zs": ev(zpp_c, zeta)[:nch], "partial_products": ev(zpp_c, zeta)[nch:],
        "quotient_polys": ev(q_c, zeta),
        "plonk_zs_next": [eval_base_poly_ext(c, zeta_next) for c in zpp_c["coeffs"][:nch]],
    }
    for key in ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products", "quotient_polys", "plonk_zs_next"):
        for e in openings[key]:
            ch.observe_ext(e)

    # ---- FRI (A.10)
    alpha = ch.get_ext_challenge()
    batches = [(zeta, [(o, j) for o in range(4) for j in range(oracles[o]["coeffs"].shape[0])]),
               (zeta_next, [(2, j) for j in range(nch)])]
    final = [E2(0)] * n
    for point, polys in batches:
        comp = [E2(0)] * n
        for (o, j) in reversed(polys):                     # sum_i alpha^i p_i
            col = oracles[o]["coeffs"][j]
            comp = [c * alpha + int(v) for c, v in zip(comp, col)]
        # divide_by_linear: (comp(X) - comp(point)) / (X - point), padded back to n coefficients
        q = [E2(0)] * n
        carry = E2(0)
        for i in range(n - 1, 0, -1):
            carry = comp[i] + carry * point
            q[i - 1] = carry
        shift = alpha.pow(len(polys))
        final = [f * shift + qq for f, qq in zip(final, q)]
    T["fri_final_poly_coeffs"] = final
    values = ext_coset_fft(final + [E2(0)] * (N - n), bits, GENERATOR)
    coeffs = final + [E2(0)] * (N - n)
    arities = fri_reduction_arity_bits(cfg, d)
    trees, fri_caps, betas_fri = [], [], []
    shift = GENERATOR
    for ab in arities:
        ar = 1 << ab
        lb = len(values).bit_length() - 1
        vr = [values[bitrev(i, lb)] for i in range(len(values))]
        leaves = np.array([[l for e in vr[i * ar:(i + 1) * ar] for l in e.limbs()] for i in range(len(vr) // ar)],
                          dtype=np.uint64)
        digests, cap = merkle_new(leaves, cfg.cap_height)
        trees.append((leaves, digests, cap))
        fri_caps.append(cap)
        ch.observe_cap(cap)
        beta = ch.get_ext_challenge()
        betas_fri.append(beta)
        new = []
        for i in range(len(coeffs) // ar):
            acc = E2(0)
            for c in reversed(coeffs[i * ar:(i + 1) * ar]):
                acc = acc * beta + c
            new.append(acc)
        coeffs = new
        shift = fpow(shift, ar)
        values = ext_coset_fft(coeffs, len(coeffs).bit_length() - 1, shift)
    final_poly = coeffs[:len(coeffs) >> rate]
    assert all(c == E2(0) for c in coeffs[len(final_poly):])
    for c in final_poly:
        ch.observe_ext(c)
    pow_witness = pow_grind(ch, cfg.proof_of_work_bits)
    ch.observe_element(pow_witness)
    ch.get_challenge()
    queries = []
    for _ in range(cfg.num_query_rounds):
        x_index = ch.get_challenge() % N
        init = [(o["leaves"][x_index].copy(), merkle_prove(o["digests"], N, cfg.cap_height, x_index)) for o in oracles]
        steps = []
        xi = x_index
        size = N
        for (leaves, digests, cap), ab in zip(trees, arities):
            ci = xi >> ab
            row = leaves[ci].copy()
            steps.append((row, merkle_prove(digests, size >> ab, cfg.cap_height, ci)))
            xi, size = ci, size >> ab
        queries.append({"x_index": x_index, "initial": init, "steps": steps})
    T.update(zeta=zeta, fri_alpha=alpha, fri_betas=betas_fri, quotient_coeffs=qcoeffs)
    return {
        "wires_cap": wires_c["cap"], "zs_pp_cap": zpp_c["cap"], "quotient_cap": q_c["cap"], "openings": openings,
        "fri_caps": fri_caps, "final_poly": final_poly, "pow_witness": pow_witness, "queries": queries,
        "public_inputs": list(public_inputs),
    }


# ------------------------------------------------------------------------------------------------ verifier
def verify(circ: Circuit, proof: dict) -> bool:
    cfg = circ.cfg
    n, d, rate = circ.n, circ.d, cfg.rate_bits
    N, bits = n << rate, d + rate
    nch = cfg.num_challenges
    pi_hash = [int(x) for x in hash_no_pad(proof["public_inputs"])]
    ch = Challenger()
    ch.observe_hash(circ.circuit_digest)
    ch.observe_hash(pi_hash)
    ch.observe_cap(proof["wires_cap"])
    betas = ch.get_n_challenges(nch)
    gammas = ch.get_n_challenges(nch)
    ch.observe_cap(proof["zs_pp_cap"])
    alphas = ch.get_n_challenges(nch)
    ch.observe_cap(proof["quotient_cap"])
    zeta = ch.get_ext_challenge()
    op = proof["openings"]
    order = ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products", "quotient_polys", "plonk_zs_next")
    for key in order:
        for e in op[key]:
            ch.observe_ext(e)
    # ---- constraint check at zeta
    npp = len(op["partial_products"]) // nch
    pps = [op["partial_products"][k * npp:(k + 1) * npp] for k in range(nch)]
    res, zh = eval_vanishing(circ, zeta, op["constants"], op["plonk_sigmas"], op["wires"], op["plonk_zs"],
                             op["plonk_zs_next"], pps, pi_hash, betas, gammas, alphas, E2(1))
    zeta_n = zeta.pow(n)
    qf = cfg.quotient_degree_factor
    for k in range(nch):
        acc = E2(0)
        for c in reversed(op["quotient_polys"][k * qf:(k + 1) * qf]):
            acc = acc * zeta_n + c
        if not (res[k] == zh * acc):
            return False
    # ---- FRI
    alpha = ch.get_ext_challenge()
    arities = fri_reduction_arity_bits(cfg, d)
    if len(proof["fri_caps"]) != len(arities):
        return False
    fri_betas = []
    for cap in proof["fri_caps"]:
        ch.observe_cap(cap)
        fri_betas.append(ch.get_ext_challenge())
    if len(proof["final_poly"]) != (n >> sum(arities)):
        return False
    for c in proof["final_poly"]:
        ch.observe_ext(c)
    if not pow_check(ch, proof["pow_witness"], cfg.proof_of_work_bits):
        return False
    ch.observe_element(proof["pow_witness"])
    ch.get_challenge()
    g_n = pyref.primitive_root_of_unity(d)
    zeta_next = zeta * g_n
    nsel_c = len(op["constants"])
    batch0 = op["constants"] + op["plonk_sigmas"] + op["wires"] + op["plonk_zs"] + op["partial_products"] + op["quotient_polys"]
    batch1 = op["plonk_zs_next"]

    def reduce_ext(vals):
        acc = E2(0)
        for v in reversed(vals):
            acc = acc * alpha + v
        return acc
    red0, red1 = reduce_ext(batch0), reduce_ext(batch1)
    caps0 = [circ.cs_commit["cap"], proof["wires_cap"], proof["zs_pp_cap"], proof["quotient_cap"]]
    wN = pyref.primitive_root_of_unity(bits)
    if len(proof["queries"]) != cfg.num_query_rounds:
        return False
    for q in proof["queries"]:
        x_index = ch.get_challenge() % N
        if q.get("x_index") is not None and x_index != q["x_index"]:      # not part of the proof bytes: derived here
            return False
        rows = []
        for (row, path), cap in zip(q["initial"], caps0):
            if not merkle_verify(row, x_index, path, cap):
                return False
            rows.append([int(v) for v in row])
        x = GENERATOR * fpow(wN, bitrev(x_index, bits)) % P
        ev0 = rows[0] + rows[1] + rows[2] + rows[3]
        ev1 = rows[2][:nch]
        s = (reduce_ext(ev0) - red0) * (E2(x) - zeta).inv()
        s = s * alpha.pow(len(ev1)) + (reduce_ext(ev1) - red1) * (E2(x) - zeta_next).inv()
        old, xi, sx = s, x_index, x
        for li, ab in enumerate(arities):
            ar = 1 << ab
            row, path = q["steps"][li]
            evals = [E2(int(row[2 * j]), int(row[2 * j + 1])) for j in range(ar)]
            ci, within = xi >> ab, xi & (ar - 1)
            if not (evals[within] == old):
                return False
            if not merkle_verify(row, ci, path, proof["fri_caps"][li]):
                return False
            # interpolate the coset {coset_start * g^i} and evaluate at beta (compute_evaluation)
            g = pyref.primitive_root_of_unity(ab)
            ev_nat = [evals[bitrev(j, ab)] for j in range(ar)]
            start = sx * fpow(g, ar - bitrev(within, ab)) % P
            pts = [start * fpow(g, j) % P for j in range(ar)]
            beta = fri_betas[li]
            tot = E2(0)
            for j in range(ar):                     # Lagrange
                num, den = E2(1), 1
                for m in range(ar):
                    if m != j:
                        num = num * (beta - pts[m])
                        den = den * (pts[j] - pts[m]) % P
                tot = tot + ev_nat[j] * num * finv(den)
            old = tot
            sx = fpow(sx, ar)
            xi = ci
        fin = E2(0)
        for c in reversed(proof["final_poly"]):
            fin = fin * sx + c
        if not (fin == old):
            return False
    return True


# ------------------------------------------------------------------------------------------------ proof bytes
def proof_bytes(proof: dict) -> bytes:
    """bincode 1.x (little-endian, fixed-width) of plonky2 v0.2.0's serde-derived `ProofWithPublicInputs`, the
    bytes the reference hex-encodes at contracts/lib/succinctx/plonky2x/core/src/utils/serde/mod.rs:82-96.
    Written field by field with struct.pack, independently of vectorx_b200/proof_io.py.  PARITY UNPINNED: the field
    order restates plonk/proof.rs and fri/proof.rs of the un-vendored crate; the reference holds no golden bytes."""
    import struct
    b = bytearray()

    def u64(v):
        b.extend(struct.pack("<Q", int(v)))

    def hash_vec(hs):
        hs = [list(h) for h in np.asarray(hs).reshape(-1, 4)]
        u64(len(hs))
        for h in hs:
            for x in h:
                u64(x)

    def ext_vec(es):
        u64(len(es))
        for e in es:
            a, c = (e.a, e.b) if isinstance(e, E2) else (e[0], e[1])
            u64(a)
            u64(c)

    hash_vec(proof["wires_cap"])
    hash_vec(proof["zs_pp_cap"])
    hash_vec(proof["quotient_cap"])
    o = proof["openings"]
    for name in ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products", "quotient_polys"):
        ext_vec(o[name])
    u64(0)                                  # lookup_zs: no lookup tables anywhere in plonky2x / VectorX
    u64(0)                                  # lookup_zs_next
    u64(len(proof["fri_caps"]))
    for cap in proof["fri_caps"]:
        hash_vec(cap)
    u64(len(proof["queries"]))
    for q in proof["queries"]:
        u64(len(q["initial"]))              # FriInitialTreeProof.evals_proofs: Vec<(Vec<F>, MerkleProof)>
        for row, path in q["initial"]:
            u64(len(row))
            for x in row:
                u64(x)
            hash_vec(path)
        u64(len(q["steps"]))                # Vec<FriQueryStep { evals: Vec<Ext>, merkle_proof }>
        for evals, path in q["steps"]:
            u64(len(evals) // 2)
            for x in evals:
                u64(x)
            hash_vec(path)
    ext_vec(proof["final_poly"])
    u64(proof["pow_witness"])
    u64(len(proof["public_inputs"]))
    for x in proof["public_inputs"]:
        u64(x)
    return bytes(b)
