"""Host-side mirror of plonky2 v0.2.0 `prove_with_partition_witness` (plonk/prover.rs), the function the
reference calls at contracts/lib/succinctx/plonky2x/core/src/backend/circuit/build.rs:69-75.

Same phase order and transcript as upstream; every heavy phase is one C-ABI call:
  wires commit            vx_commit_from_values          ("FFT + blinding", "build Merkle tree")
  Z / partial products    vx_zs_partial_products
  quotient                vx_quotient + vx_commit_from_coeffs  ("compute quotient polys")
  openings                vx_batch_eval_ext
  FRI                     vx_fri_begin / commit_layer / fold / final_poly / pow_grind / query
The challenger and the proof assembly stay on the host.
"""
from __future__ import annotations

import ctypes
import time

import numpy as np

from . import gates as gate_lib
from ._lib import Context, DeviceArray, VxError, check, default_context, load, ptr, vp
from .challenger import Challenger, hash_no_pad_host, hash_pad_host
from .plonky2 import PolynomialBatch

P = 0xFFFFFFFF00000001
GENERATOR = 14293326489335486720


class VxCircuitDesc(ctypes.Structure):
    _fields_ = [("degree_bits", ctypes.c_uint32), ("rate_bits", ctypes.c_uint32), ("num_wires", ctypes.c_uint32),
                ("num_routed_wires", ctypes.c_uint32), ("num_constants", ctypes.c_uint32),
                ("num_selectors", ctypes.c_uint32), ("num_challenges", ctypes.c_uint32),
                ("num_partial_products", ctypes.c_uint32), ("max_degree", ctypes.c_uint32),
                ("num_gate_constraints", ctypes.c_uint32), ("k_is", ctypes.c_void_p), ("program", ctypes.c_void_p),
                ("program_len", ctypes.c_uint64)]


class FriRange(ctypes.Structure):
    _fields_ = [("oracle", ctypes.c_uint32), ("first", ctypes.c_uint32), ("count", ctypes.c_uint32)]


class FriBatch(ctypes.Structure):
    _fields_ = [("point", ctypes.c_uint64 * 2), ("ranges", ctypes.POINTER(FriRange)), ("num_ranges", ctypes.c_uint32)]


class CircuitData:
    """CommonCircuitData + the prover-only parts this path needs (plain data, as the Rust shim would pass it)."""

    def __init__(self, degree_bits, gate_ids, selector_index, groups, constants, sigmas, *, num_wires=135,
                 num_routed_wires=80, rate_bits=3, cap_height=4, num_challenges=2, max_degree=8,
                 quotient_degree_factor=8, num_query_rounds=28, proof_of_work_bits=16, arity_bits=4,
                 final_poly_bits=5, ctx: Context | None = None, superops: bool = True,
                 domain_separator=(), domain_separator_digest=None, circuit_digest=None, compile_gates: bool = False):
        # plonky2 keeps three independent parameters -- max_quotient_degree_factor (= max_degree here), the circuit's
        # quotient_degree_factor (chunk size of the permutation argument, number of quotient chunks) and 2^rate_bits (size
        # of the LDE the quotient is evaluated on).  The device path chunks by max_degree and reshapes the quotient as
        # 2^rate_bits chunks, which is only the same thing when all three agree (standard_recursion_config: 8 / 8 / 3,
        # contracts/lib/succinctx/plonky2x/core/src/frontend/builder/mod.rs:69): anything else is rejected, not mis-proved.
        if not (quotient_degree_factor == max_degree == (1 << rate_bits)):
            raise ValueError(f"unsupported circuit config: quotient_degree_factor={quotient_degree_factor}, "
                             f"max_degree={max_degree}, 2^rate_bits={1 << rate_bits} must all be equal")
        self.ctx = ctx or default_context()
        self.degree_bits, self.n = degree_bits, 1 << degree_bits
        self.gate_ids, self.selector_index, self.groups = list(gate_ids), list(selector_index), [tuple(g) for g in groups]
        self.num_selectors = len(self.groups)
        self.num_wires, self.num_routed_wires = num_wires, num_routed_wires
        self.rate_bits, self.cap_height, self.num_challenges = rate_bits, cap_height, num_challenges
        self.max_degree, self.quotient_degree_factor = max_degree, quotient_degree_factor
        self.num_query_rounds, self.proof_of_work_bits = num_query_rounds, proof_of_work_bits
        self.arity_bits, self.final_poly_bits = arity_bits, final_poly_bits
        self.constants = np.ascontiguousarray(constants, dtype=np.uint64)
        self.sigmas = np.ascontiguousarray(sigmas, dtype=np.uint64)
        self.sigmas_dev = DeviceArray.from_host(self.ctx, self.sigmas)      # prover-only data: resident for every proof
        self.num_constants = self.constants.shape[0]
        self.num_partial_products = (num_routed_wires + max_degree - 1) // max_degree - 1
        self.k_is = np.array([pow(GENERATOR, j, P) for j in range(num_routed_wires)], dtype=np.uint64)
        meta = [gate_lib.lookup(g) for g in self.gate_ids]
        self.num_gate_constraints = max(m[4] for m in meta)
        self.program = gate_lib.build_program(self.gate_ids, self.selector_index, self.groups, self.num_selectors,
                                              superops=superops)
        # CircuitBuilder::build: constants_sigmas commitment and the circuit digest
        cs = np.concatenate([self.constants, self.sigmas])
        self.constants_sigmas_commitment = PolynomialBatch.from_values(cs, rate_bits, False, cap_height, ctx=self.ctx)
        cap = self.constants_sigmas_commitment.cap.hashes
        # plonky2 `CircuitBuilder::build`: circuit_digest = hash_no_pad(cap.flatten() ++ hash_pad(domain_separator) ++
        # [degree_bits]); the domain separator is empty unless the builder set one.  A caller that holds the Rust-side
        # values (CommonCircuitData / VerifierOnlyCircuitData) passes them in and they are used verbatim.
        if domain_separator_digest is None:
            domain_separator_digest = hash_pad_host(list(domain_separator))
        self.domain_separator_digest = [int(x) % P for x in domain_separator_digest]
        if circuit_digest is None:
            circuit_digest = hash_no_pad_host([int(x) for x in cap.reshape(-1)] + self.domain_separator_digest + [degree_bits])
        self.circuit_digest = [int(x) % P for x in circuit_digest]
        self.desc = VxCircuitDesc(degree_bits, rate_bits, num_wires, num_routed_wires, self.num_constants,
                                  self.num_selectors, num_challenges, self.num_partial_products, max_degree,
                                  self.num_gate_constraints, self.k_is.ctypes.data, self.program.ctypes.data,
                                  len(self.program))

        # circuit-load-time specialisation: the gate program compiled to a straight-line sm_100a kernel (NVRTC, tens of
        # seconds once per circuit); vx_quotient then runs it instead of interpreting the bytecode.  Same results.
        self.gates_compiled, self.gates_compile_error = False, None
        if compile_gates:
            self.compile_gates(strict=False)

    def compile_gates(self, tuning: int = 0, strict: bool = True) -> bool:
        """tuning: 0 = defaults, else min blocks per SM | (threads per block / 32) << 8 | (operations per fence) << 16.
        strict=False: a box without NVRTC (or a failed compilation) leaves the circuit on the bytecode interpreter -- the
        same GPU path it used before, not a fallback to the CPU -- and the reason in `gates_compile_error`."""
        rc = load().vx_quotient_compile(self.ctx.handle, ctypes.byref(self.desc), tuning)
        if rc != 0:
            self.gates_compile_error = f"code {rc}: " + load().vx_last_error().decode("utf-8", "replace")
            if strict:
                raise VxError("vx_quotient_compile failed with " + self.gates_compile_error)
        self.gates_compiled = bool(load().vx_quotient_is_compiled(self.ctx.handle, ctypes.byref(self.desc)))
        return self.gates_compiled

    def close(self):
        """Release the device-resident prover data (before the context it lives on is destroyed)."""
        self.constants_sigmas_commitment.close()
        self.sigmas_dev.close()

    def fri_reduction_arity_bits(self):
        out, d = [], self.degree_bits
        while d > self.final_poly_bits and d + self.rate_bits - self.arity_bits >= self.cap_height:
            out.append(self.arity_bits)
            d -= self.arity_bits
        return out


def _u64(xs):
    return np.ascontiguousarray(np.array([int(x) % P for x in xs], dtype=np.uint64))


def prove(circ: CircuitData, wires: np.ndarray, public_inputs, trace: dict | None = None) -> dict:
    """prove_with_partition_witness: wires is the (num_wires, n) witness matrix. Returns the proof as a dict."""
    lib, ctx = load(), circ.ctx
    T = trace if trace is not None else {}
    phases = T.setdefault("phase_ms", {})       # plonky2's TimingTree scopes: every C call is blocking, so wall time is device time + launch
    _t = [time.perf_counter()]

    def lap(name):
        now = time.perf_counter()
        phases[name] = phases.get(name, 0.0) + (now - _t[0]) * 1e3
        _t[0] = now
    n, d, rate, cap_h = circ.n, circ.degree_bits, circ.rate_bits, circ.cap_height
    N, bits, nch = n << rate, d + rate, circ.num_challenges
    want_trace = trace is not None and trace.get("intermediates", True)
    wires = np.ascontiguousarray(wires, dtype=np.uint64)
    pi_hash = hash_no_pad_host(list(public_inputs))
    ch = Challenger()
    ch.observe_hash(circ.circuit_digest)
    ch.observe_hash(pi_hash)
    lap("host")
    # the witness goes to the GPU once, inside the commit's own column pipeline (copy of chunk k+1 under the transforms
    # and hashing of chunk k); every later phase reads the device copy
    wires_d = DeviceArray(ctx, wires.shape)
    wires_c = PolynomialBatch.from_values_keep(wires, wires_d, rate, cap_h, ctx=ctx)
    lap("commit wires")
    ch.observe_cap(wires_c.cap.hashes.tolist())
    betas = ch.get_n_challenges(nch)
    gammas = ch.get_n_challenges(nch)
    zpp = DeviceArray(ctx, (nch * (1 + circ.num_partial_products), n))
    a_betas, a_gammas = _u64(betas), _u64(gammas)        # keep the arrays alive across the FFI calls
    check(lib.vx_zs_partial_products(ctx.handle, ctypes.byref(circ.desc), ptr(wires_d), ptr(circ.sigmas_dev),
                                     ptr(a_betas), ptr(a_gammas), ptr(zpp)), "vx_zs_partial_products")
    lap("Z / partial products")
    zpp_c = PolynomialBatch.from_values(zpp, rate, False, cap_h, ctx=ctx)
    lap("commit Z / partial products")
    ch.observe_cap(zpp_c.cap.hashes.tolist())
    alphas = ch.get_n_challenges(nch)
    qcoeffs = DeviceArray(ctx, (nch, N))
    cs_c = circ.constants_sigmas_commitment
    a_pi, a_alphas = _u64(pi_hash), _u64(alphas)
    check(lib.vx_quotient(ctx.handle, ctypes.byref(circ.desc), cs_c.handle, wires_c.handle, zpp_c.handle,
                          ptr(a_pi), ptr(a_betas), ptr(a_gammas), ptr(a_alphas), ptr(qcoeffs)), "vx_quotient")
    lap("compute quotient polys")
    qchunks = qcoeffs.reshape(nch * circ.quotient_degree_factor, n)
    q_c = PolynomialBatch.from_coeffs(qchunks, rate, False, cap_h, ctx=ctx)
    lap("commit quotient")
    ch.observe_cap(q_c.cap.hashes.tolist())
    zeta = ch.get_extension_challenge()
    g_n = pow(7277203076849721926, 1 << (32 - d), P)
    zeta_next = [zeta[0] * g_n % P, zeta[1] * g_n % P]
    T.update(betas=betas, gammas=gammas, alphas=alphas, zeta=zeta, pi_hash=pi_hash)
    if want_trace:                                       # tests compare the intermediates; a proof never needs them on the host
        T.update(zpp=zpp.to_host(), quotient_coeffs=qchunks.to_host())
    for buf in (wires_d, zpp, qcoeffs):
        buf.close()

    def ev(batch, point):
        out = np.zeros((batch.num_polys, 2), dtype=np.uint64)
        a_point = _u64(point)
        check(lib.vx_batch_eval_ext(batch.handle, ptr(a_point), ptr(out)), "vx_batch_eval_ext")
        return [[int(a), int(b)] for a, b in out]
    cs_open, z_open = ev(cs_c, zeta), ev(zpp_c, zeta)
    openings = {
        "constants": cs_open[:circ.num_constants], "plonk_sigmas": cs_open[circ.num_constants:],
        "wires": ev(wires_c, zeta), "plonk_zs": z_open[:nch], "partial_products": z_open[nch:],
        "quotient_polys": ev(q_c, zeta), "plonk_zs_next": ev(zpp_c, zeta_next)[:nch],
    }
    lap("openings")
    for key in ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products", "quotient_polys", "plonk_zs_next"):
        for e in openings[key]:
            ch.observe_extension_element(e)

    # ---- FRI
    alpha = ch.get_extension_challenge()
    oracles = [cs_c, wires_c, zpp_c, q_c]
    handles = (ctypes.c_void_p * 4)(*[o.handle.value for o in oracles])
    r0 = (FriRange * 4)(*[FriRange(i, 0, o.num_polys) for i, o in enumerate(oracles)])
    r1 = (FriRange * 1)(FriRange(2, 0, nch))
    batches = (FriBatch * 2)()
    batches[0].point[0], batches[0].point[1] = zeta
    batches[0].ranges, batches[0].num_ranges = r0, 4
    batches[1].point[0], batches[1].point[1] = zeta_next
    batches[1].ranges, batches[1].num_ranges = r1, 1
    fri = vp()
    a_alpha = _u64(alpha)
    lap("host")
    check(lib.vx_fri_begin(ctx.handle, handles, 4, batches, 2, ptr(a_alpha), ctypes.byref(fri)), "vx_fri_begin")
    lap("FRI batch + LDE")
    try:
        arities = circ.fri_reduction_arity_bits()
        fri_caps = []
        for ab in arities:
            cap = np.zeros((1 << cap_h, 4), dtype=np.uint64)
            check(lib.vx_fri_commit_layer(fri, ab, cap_h, ptr(cap)), "vx_fri_commit_layer")
            fri_caps.append(cap)
            ch.observe_cap(cap.tolist())
            a_beta = _u64(ch.get_extension_challenge())
            check(lib.vx_fri_fold(fri, ptr(a_beta)), "vx_fri_fold")
        flen = ctypes.c_uint32(0)
        fbuf = np.zeros((n, 2), dtype=np.uint64)
        check(lib.vx_fri_final_poly(fri, ptr(fbuf), ctypes.byref(flen)), "vx_fri_final_poly")
        final_poly = [[int(a), int(b)] for a, b in fbuf[:flen.value]]
        for c in final_poly:
            ch.observe_extension_element(c)
        lap("FRI fold-and-commit")
        st, pos = ch.pow_state()
        wit = ctypes.c_uint64(0)
        a_st = _u64(st)
        check(lib.vx_pow_grind(ctx.handle, ptr(a_st), pos, circ.proof_of_work_bits, ctypes.byref(wit)), "vx_pow_grind")
        lap("FRI proof of work")
        pow_witness = int(wit.value)
        ch.observe_element(pow_witness)
        ch.get_challenge()
        x_indices = [ch.get_challenge() % N for _ in range(circ.num_query_rounds)]
        k = len(x_indices)
        init_rows = [o.leaves(x_indices) for o in oracles]
        init_paths = [o.prove(x_indices) for o in oracles]
        layer_rows, layer_paths = [], []
        idx = list(x_indices)
        size = N
        for li, ab in enumerate(arities):
            idx = [x >> ab for x in idx]
            size >>= ab
            depth = max((size.bit_length() - 1) - cap_h, 0)
            rows = np.zeros((k, 2 << ab), dtype=np.uint64)
            paths = np.zeros((k, max(depth, 1), 4), dtype=np.uint64)
            a_idx = _u64(idx)
            check(lib.vx_fri_query(fri, li, ptr(a_idx), k, ptr(rows), ptr(paths)), "vx_fri_query")
            layer_rows.append(rows)
            layer_paths.append(paths[:, :depth])
        lap("FRI queries")
    finally:
        lib.vx_fri_free(fri)
    queries = []
    for qi, x in enumerate(x_indices):
        queries.append({"x_index": x,
                        "initial": [(init_rows[o][qi], init_paths[o][qi]) for o in range(4)],
                        "steps": [(layer_rows[li][qi], layer_paths[li][qi]) for li in range(len(arities))]})
    proof = {"wires_cap": wires_c.cap.hashes, "zs_pp_cap": zpp_c.cap.hashes, "quotient_cap": q_c.cap.hashes,
             "openings": openings, "fri_caps": fri_caps, "final_poly": final_poly, "pow_witness": pow_witness,
             "queries": queries, "public_inputs": list(public_inputs)}
    for b in (wires_c, zpp_c, q_c):
        b.close()
    lap("host")
    return proof
