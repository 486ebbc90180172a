#!/usr/bin/env python
"""bench.py -- LDE + Poseidon-Merkle commit throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one plonky2 `PolynomialBatch::from_values` commit (iNTT -> coset LDE -> Poseidon Merkle
tree to the cap) of the workload shape, through the C ABI of include/vectorx_b200.h.
 * value : whole-job Melem/s (input trace elements n*c per second) with the values already in HBM;
 * e2e   : the same through the public API with pinned HOST buffers: H2D of the values and D2H of
           the cap inside the timed region;
 * N > 1 : ONE commit is sharded over the ranks (strong scaling): column-sharded iNTT, NCCL
           all-gather of the coefficients, each rank extends and hashes its own cosets / cap
           subtrees (coset partition), caps gathered with NCCL;
 * --impl reference : the CPU restatement of the reference path (oracle/, OpenMP over columns /
           subtrees like plonky2's Rayon) timed on the host cores.  The real Rust prover cannot be
           built here (no cargo; plonky2 not vendored) -- kind = "port".
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P = 0xFFFFFFFF00000001
METRIC = "LDE+Poseidon-Merkle commit throughput"
UNIT = "Melem/s"
IMADS_PER_PERMUTATION = 6612        # 1077 full mults x 4 + 1152 small MACs x 2
N_INPUT_SETS = 4                     # rotate inputs so consecutive steps never hit L2 (4 x 70.8 MB > 126 MB)


def workload_from_args(a):
    return dict(log_n=a.log_n, cols=a.cols, rate_bits=a.rate_bits, cap_height=a.cap_height)


def workload_name(w):
    return f"plonky2 from_values commit 2^{w['log_n']} rows x {w['cols']} cols, rate_bits={w['rate_bits']}, cap_height={w['cap_height']}"


def gen_values(c, n, seed, col_lo=0, col_hi=None):
    """uniform canonical field elements; column j depends on (seed, j) only, so a rank can generate just its own slice"""
    col_hi = c if col_hi is None else col_hi
    a = np.empty((col_hi - col_lo, n), dtype=np.uint64)
    for j in range(col_lo, col_hi):
        a[j - col_lo] = np.random.default_rng([seed, j]).integers(0, P, size=n, dtype=np.uint64)
    return a


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_int_peak():
    """thread-level 32-bit IMAD issue rate measured by tools/intpeak on this pool's B200 (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "int_peaks_r01.json")) as f:
            d = json.load(f)
        return float(d["imad_lo"]["thread_instr_per_s"]) / 1e9, "measured IMAD issue rate (profiles/int_peaks_r01.json)"
    except Exception:
        return 148 * 64 * 1.965, "nominal 148 SM x 64 IMAD/clk x 1.965 GHz"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_commit_seconds(w, log_n_sample, repeats=1):
    """Times the oracle (native build, all host threads) on a bounded sample of the workload."""
    import oracle                                   # the ONLY place bench.py touches oracle/: the CPU baseline
    c = w["cols"]
    cols = gen_values(c, 1 << log_n_sample, seed=99)
    L = oracle.lib(native=True)
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to every rank, which would time the
    # CPU arm single-threaded -- ask the scheduler instead of the environment
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    L.vxo_set_num_threads(int(avail))
    threads = L.vxo_num_threads()
    best = None
    for _ in range(repeats):
        t = time.perf_counter()
        oracle.commit_from_values(cols, w["rate_bits"], w["cap_height"], want_leaves=True, want_digests=True, native=True)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return best, threads, c * (1 << log_n_sample)


def run_reference(a, w, rank):
    """--impl reference: the reference's CPU path (port) on host cores; rank 0 only.  Every step is one FULL commit of
    the configured shape (same config as the GPU arm); only shapes whose CPU commit would not fit the time budget
    (> 2^17 rows x 135 cols worth of elements) fall back to a row sample, and then the workload name says so."""
    if rank != 0:
        return
    budget_elems = 135 << 17
    sample_log_n = w["log_n"]
    while sample_log_n > 10 and (w["cols"] << sample_log_n) > budget_elems:
        sample_log_n -= 1
    full = sample_log_n == w["log_n"]
    for _ in range(min(max(a.warmup, 1), 2)):                    # CPU needs no long warm-up; bound the run
        cpu_commit_seconds(w, sample_log_n)
    times = []
    elems = 0
    threads = 1
    for _ in range(a.steps):
        dt, threads, elems = cpu_commit_seconds(w, sample_log_n)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = elems / (ms * 1e-3) / 1e6
    ws = dict(w, log_n=sample_log_n)
    sample = ("one full commit of the configured shape per step" if full else
              f"2^{sample_log_n} rows x {w['cols']} cols (1/{1 << (w['log_n'] - sample_log_n)} of the rows), same rate_bits/cap")
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64 (Goldilocks)", "data": "synthetic", "same_config": full,
            "config": {"workload": workload_name(w) if full else workload_name(ws) + f" (row sample of 2^{w['log_n']})",
                       "note": "CPU restatement of plonky2 v0.2.0 (oracle/), OpenMP; "
                       "the Rust reference is unbuildable here (no cargo, plonky2 not vendored)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ whole proofs
def run_prove_block(a, vx, ctx, dist, rank, local_rank, G):
    """Second half of BASELINE's metric ("prove secs at 1/2/4/8 GPU").  The header_range / rotate circuits need the Rust
    witness generator (not buildable here), so the proof is of a synthetic circuit with the reference's config
    (standard_recursion_config: 135 wires / 80 routed, rate_bits 3, cap_height 4, 28 queries, 16-bit PoW) and a 9-gate mix,
    2^prove_bits rows.  Circuit, witness and the verifier come from oracle/ (test infrastructure: input generation and the
    untimed check only); everything timed goes through the C ABI.
      ms_per_proof : one prove_with_partition_witness from a pageable witness on one GPU (best of 3, rank 0);
      proofs_per_s : LocalProver.batch_prove (P2X/backend/prover/local.rs:34-48) -- every rank proves `prove_batch`
                     independent inputs on its own GPU with four workers sharing one context (one per lane); all proofs / max time."""
    import torch
    from oracle import plonk, synth                      # inputs + untimed verification only
    from oracle.field import E2
    from vectorx_b200.local_prover import CircuitSpec, LocalProver
    bits = a.prove_bits
    circ, wires, pis = synth.build(bits, seed=11)
    # compile_gates: the circuit's gate program is compiled once at load time (vx_quotient_compile, untimed like the
    # circuit build itself); --no-compile-gates interprets the bytecode instead
    spec = CircuitSpec(circ.d, [g.id() for g in circ.gates], circ.selector_index, circ.groups, circ.constants, circ.sigmas,
                       compile_gates=not a.no_compile_gates)
    lp = LocalProver(devices=[local_rank], workers_per_device=4)
    lp.batch_prove(spec, [(wires, pis)] * 8)             # warm-up: circuit replica, every lane's pools and staging buffers
    bound = [c for r in lp._replicas if r is not None for c in r.bound.values()]
    compiled = bool(bound) and all(getattr(c, "gates_compiled", False) for c in bound)
    compile_err = next((c.gates_compile_error for c in bound if getattr(c, "gates_compile_error", None)), None)
    ms_single, verified = None, None
    if rank == 0:
        runs = []
        for _ in range(3):
            t = time.perf_counter()
            proof = lp.prove(spec, (wires, pis))
            runs.append((time.perf_counter() - t) * 1e3)
        ms_single = min(runs)
        oproof = dict(proof)
        oproof["openings"] = {k: [E2(int(e[0]), int(e[1])) for e in v] for k, v in proof["openings"].items()}
        oproof["final_poly"] = [E2(int(e[0]), int(e[1])) for e in proof["final_poly"]]
        verified = bool(plonk.verify(circ, oproof))
        if not verified:
            raise SystemExit("bench.py: the GPU proof was rejected by the oracle verifier")
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    proofs = lp.batch_prove(spec, [(wires, pis)] * a.prove_batch)
    dt = time.perf_counter() - t
    ref_w = proofs[0]["pow_witness"]
    same = all(p["pow_witness"] == ref_w for p in proofs)
    tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    lp.close()
    if rank != 0:
        return None
    return {"circuit": f"synthetic 2^{bits}-row circuit, standard_recursion_config, gates: " + ", ".join(sorted({g.id().split('{')[0].split('(')[0].split(' ')[0] for g in circ.gates})),
            "witness": "pageable host memory", "ms_per_proof": ms_single, "secs_per_proof": ms_single / 1e3,
            "proofs_per_s": G * a.prove_batch / float(tt.item()), "proofs_per_batch_per_gpu": a.prove_batch, "n_gpus": G,
            "workers_per_gpu": 4, "verified_by_oracle_verifier": verified, "batch_identical": same,
            "gate_program": "compiled at circuit load (NVRTC, sm_100a)" if compiled else
                            "interpreted bytecode" + (f" (compilation failed: {compile_err})" if compile_err else ""),
            "note": "BASELINE configs 2-4 (header_range_256/512, rotate) need the Rust witness generator: not measured here"}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=16)
    ap.add_argument("--cols", type=int, default=135)
    ap.add_argument("--rate-bits", type=int, default=3)
    ap.add_argument("--cap-height", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prove", action="store_true", help="skip the whole-proof block (second half of BASELINE's metric)")
    ap.add_argument("--no-compile-gates", action="store_true", help="prove block: interpret the gate bytecode (no NVRTC)")
    ap.add_argument("--prove-bits", type=int, default=16, help="rows (log2) of the synthetic circuit of the prove block")
    ap.add_argument("--prove-batch", type=int, default=32, help="proofs per rank in the batch_prove throughput run")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: 'peer' = the library's own kernels over NVLink peer memory, 'nccl' = torch.distributed all-gather (A/B)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    w = workload_from_args(a)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        run_reference(a, w, rank)
        return

    import torch
    import vectorx_b200 as vx
    from vectorx_b200._lib import check, load, ptr, vp
    import ctypes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no GPU visible; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    G = world
    ctx = vx.Context(local_rank)
    lib = load()
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    c, log_n, rate, cap = w["cols"], w["log_n"], w["rate_bits"], w["cap_height"]
    n = 1 << log_n
    N = n << rate
    elems = n * c
    if G > 1:
        assert G <= (1 << cap), "sharding needs whole cap subtrees per rank: G <= 2^cap_height"
    cpr = (c + G - 1) // G                     # columns per rank for the iNTT stage (last rank zero-padded)
    col_lo = rank * cpr
    col_hi = min(c, col_lo + cpr)

    # ---- which global columns this rank transforms (contiguous slice, or the library's interleaved parts for big slices)
    col_map = [j if j < c else -1 for j in range(col_lo, col_lo + cpr)]
    group = None
    if G > 1 and a.exchange == "peer":
        from vectorx_b200.sharded import PeerGroup, ShardPlan
        group = PeerGroup(ctx, ShardPlan(G, rank, c, log_n, rate, cap), dist)
        col_map = group.column_map()
    # ---- inputs: N_INPUT_SETS different value matrices, pinned on the host and resident in HBM
    host_sets, dev_sets = [], []
    for s in range(N_INPUT_SETS):
        mine = np.zeros((cpr, n), dtype=np.uint64)
        for i, gcol in enumerate(col_map):
            if gcol >= 0:
                mine[i] = gen_values(c, n, 0x5EED0001 + s, gcol, gcol + 1)[0]
        h = torch.from_numpy(mine.view(np.int64)).pin_memory()
        host_sets.append(h)
        dev_sets.append(h.to(f"cuda:{local_rank}"))
    # the same values as plonky2 holds them: c separately allocated PAGEABLE columns (Vec<PolynomialValues<F>>)
    page_sets = []
    if G == 1:
        for s_ in range(N_INPUT_SETS):
            full = host_sets[s_].numpy().view(np.uint64)
            page_sets.append([np.array(full[j], copy=True) for j in range(c)])
    coeff_all = torch.empty((G * cpr, n), dtype=torch.int64, device=f"cuda:{local_rank}")
    coeff_mine = torch.empty((cpr, n), dtype=torch.int64, device=f"cuda:{local_rank}")
    caps_loc = (1 << cap) // G
    cap_loc = torch.empty((caps_loc, 4), dtype=torch.int64, device=f"cuda:{local_rank}")
    cap_all = torch.empty((1 << cap, 4), dtype=torch.int64, device=f"cuda:{local_rank}")
    cap_host = torch.empty((1 << cap, 4), dtype=torch.int64).pin_memory()
    phase_acc = {}
    use_nccl = False
    if G > 1:
        from vectorx_b200.sharded import DeviceEngine, PeerGroup, ShardPlan, TorchComm, sharded_commit
        plan = ShardPlan(G, rank, c, log_n, rate, cap)
        use_nccl = a.exchange == "nccl"
        if use_nccl:
            engine, comm = DeviceEngine(ctx), TorchComm(dist)
            bufs = {"coeff_mine": coeff_mine, "coeff_all": coeff_all, "cap_loc": cap_loc, "cap_all": cap_all}

    def one_step(src, from_host):
        """One commit. src: this rank's (cpr, n) values (host-pinned or device). Returns nothing; cap -> cap_host."""
        if G == 1:
            h = vp()
            check(lib.vx_commit_from_values(ctx.handle, src.data_ptr(), c, log_n, rate, cap, ctypes.byref(h)),
                  "vx_commit_from_values")
            for k, v in ctx.phase_ms().items():
                phase_acc[k] = phase_acc.get(k, 0.0) + v
            if from_host:
                check(lib.vx_batch_cap(h, cap_host.data_ptr()), "vx_batch_cap")
            lib.vx_batch_free(h)
            return
        if not use_nccl:
            # the library's own exchange: coefficient blocks stored into every rank's gather buffer over NVLink peer
            # memory, flags instead of collectives (csrc/shard.cu); one blocking call per rank
            h = group.commit_from_values(src, cap_host if from_host else cap_all)
            for k, v in ctx.phase_ms().items():
                phase_acc[k] = phase_acc.get(k, 0.0) + v
            group.free_batch(h)
            return
        # column-sharded iNTT -> NCCL all-gather of coefficients -> own cosets / cap subtrees -> caps gathered
        h, _ = sharded_commit(src, plan, engine, comm, bufs)
        for k, v in ctx.phase_ms().items():
            phase_acc[k] = phase_acc.get(k, 0.0) + v
        if from_host and rank == 0:
            cap_host.copy_(cap_all, non_blocking=False)
        engine.free(h)

    def one_step_cols(cols):
        """G == 1: one commit through vx_commit_from_values_cols from c separate pageable allocations; cap -> cap_host."""
        h = vp()
        ptrs = (vp * c)(*[x.ctypes.data for x in cols])
        check(lib.vx_commit_from_values_cols(ctx.handle, ptrs, c, log_n, rate, cap, ctypes.byref(h)),
              "vx_commit_from_values_cols")
        check(lib.vx_batch_cap(h, cap_host.data_ptr()), "vx_batch_cap")
        lib.vx_batch_free(h)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def timed(from_host, steps, per_column=False):
        sets = host_sets if from_host else dev_sets
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # G == 1: events on the library's stream.  G > 1: the NCCL legs run on torch's current stream and every library
        # call is blocking, so events on torch's stream bracket all device work of the region.
        ev_stream = stream if (G == 1 or not use_nccl) else torch.cuda.current_stream()
        t0 = time.perf_counter()
        e0.record(ev_stream)
        for i in range(steps):
            if per_column:
                one_step_cols(page_sets[i % N_INPUT_SETS])
            else:
                one_step(sets[i % N_INPUT_SETS], from_host)
        e1.record(ev_stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall

    # ---- warm-up, then the two timed regions
    for i in range(a.warmup):
        one_step(dev_sets[i % N_INPUT_SETS], False)
    for i in range(2):
        one_step(host_sets[i % N_INPUT_SETS], True)
    shard_parity = None
    if G > 1:
        # parity (untimed), on EVERY rank: a whole single-GPU commit of the same values on this rank's own GPU, then
        #  * this rank's block of Merkle digests (its cosets' leaf digests and interior levels, plonky2 layout) must be
        #    the matching slice of the whole commit's digests,
        #  * the gathered cap must equal the whole cap -- for device-resident and for host values (streamed pipeline).
        full0 = gen_values(c, n, seed=0x5EED0001)
        whole = vx.PolynomialBatch.from_values(full0, rate, False, cap, ctx=ctx)
        whole_digests = np.zeros((2 * (N - (1 << cap)), 4), dtype=np.uint64)
        check(lib.vx_batch_download(whole._h, None, whole_digests.ctypes.data), "vx_batch_download")
        per_rank = whole_digests.shape[0] // G
        ok = True
        for from_host in (False, True):
            src = host_sets[0] if from_host else dev_sets[0]
            if use_nccl:
                h, _ = sharded_commit(src, plan, engine, comm, bufs)
                got_cap = cap_all.cpu().numpy().view(np.uint64)
                free = engine.free
            else:
                h = group.commit_from_values(src, cap_host if from_host else cap_all)
                got_cap = (cap_host.numpy() if from_host else cap_all.cpu().numpy()).view(np.uint64)
                free = group.free_batch
            torch.cuda.synchronize()
            mine = np.zeros((per_rank, 4), dtype=np.uint64)
            check(lib.vx_batch_download(h, None, mine.ctypes.data), "vx_batch_download")
            ok = ok and np.array_equal(mine, whole_digests[rank * per_rank:(rank + 1) * per_rank])
            ok = ok and (use_nccl and from_host and rank != 0 or np.array_equal(got_cap, whole.cap.hashes))
            free(h)
        whole.close()
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=f"cuda:{local_rank}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        shard_parity = bool(flag.item())
        if not shard_parity:
            raise SystemExit("bench.py: a rank's digests / the gathered cap differ from the single-GPU commit")
        del full0, whole_digests
    launches0 = ctx.launch_count
    phase_acc.clear()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, wall_ms = timed(False, a.steps)
    launches = ctx.launch_count - launches0
    phases = {k: v / a.steps for k, v in phase_acc.items()}
    e2e_ms, _ = timed(True, a.steps)
    e2e_cols_ms = None
    if G == 1:
        for i in range(2):
            one_step_cols(page_sets[i])
        e2e_cols_ms, _ = timed(True, a.steps, per_column=True)
    clocks = sampler.stop() if rank == 0 else None
    prove_block = None if a.no_prove else run_prove_block(a, vx, ctx, dist, rank, local_rank, G)

    ms_per_step = total_ms / a.steps
    value = elems / (ms_per_step * 1e-3) / 1e6
    e2e_value = elems / (e2e_ms / a.steps * 1e-3) / 1e6

    if dist is not None:
        lt = torch.tensor([launches], dtype=torch.int64, device=f"cuda:{local_rank}")
        dist.all_reduce(lt)
        launches = int(lt.item())

    if rank == 0:
        hbm_peak, hbm_src = load_peaks()
        int_peak, int_src = load_int_peak()
        # dominant kernel: leaf hashing (one launch per commit per rank)
        rows_loc = N // G
        leaf_ms = phases.get("leaf_hash", float("nan"))
        perms = rows_loc * ((c + 7) // 8) if c > 4 else 0
        leaf_bytes = 8 * rows_loc * c + 32 * rows_loc                 # read the LDE rows once, write digests
        leaf_gops = perms * IMADS_PER_PERMUTATION / (leaf_ms * 1e-3) / 1e9 if leaf_ms > 0 else None
        leaf_gbs = leaf_bytes / (leaf_ms * 1e-3) / 1e9 if leaf_ms > 0 else None
        ntt_ms = phases.get("intt", 0.0) + phases.get("lde", 0.0)
        ntt_bytes = (8 * n * c * 2 if G == 1 else 0) + 8 * n * c + 8 * rows_loc * c   # values->coeffs, coeffs->LDE
        traffic = None
        try:                                # dram read+write bytes of this kernel from the newest committed ncu --set full digest
            import glob
            cand = sorted(glob.glob(os.path.join(ROOT, "profiles", "*leaf_hash_ncu_traffic.json")))
            with open(cand[-1]) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        pipes = None
        try:
            with open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r0*_leaf_hash_pipes.json")))[-1]) as f:
                pipes = json.load(f)
        except Exception:
            pass
        wide_peak = None
        try:
            with open(os.path.join(ROOT, "profiles", "int_peaks_r01.json")) as f:
                wide_peak = float(json.load(f)["imad_wide_acc"]["thread_instr_per_s"]) / 1e9
        except Exception:
            pass
        roofline = {
            "kernel": "leaf_hash_kernel<col_major> (Poseidon sponge, one thread per LDE row)",
            "bound": "int_alu",
            "achieved": leaf_gops, "peak": int_peak, "unit": "G 32x32 mul-add/s",
            "frac": (leaf_gops / int_peak) if leaf_gops else None,
            "peak_source": int_src,
            # the counted unit (32x32+64 multiply-add) is one IMAD.WIDE.U32, whose own measured issue rate is lower than
            # the 32-bit IMAD rate used as `peak`; both fractions are reported, plus what ncu saw on the pipes
            "peak_imad_wide": wide_peak,
            "frac_vs_imad_wide": (leaf_gops / wide_peak) if (leaf_gops and wide_peak) else None,
            "ncu_pipes": pipes,
            "algorithmic_ops_per_launch": perms * IMADS_PER_PERMUTATION,
            "ms_per_launch": leaf_ms,
            "traffic": traffic,
            "hbm": {"bound": "hbm", "achieved": leaf_gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (leaf_gbs / hbm_peak) if leaf_gbs else None, "algorithmic_bytes_per_launch": leaf_bytes,
                    "peak_source": hbm_src},
            "ntt_passes": {"bound": "hbm", "ms": ntt_ms, "algorithmic_bytes": ntt_bytes,
                           "achieved": ntt_bytes / (ntt_ms * 1e-3) / 1e9 if ntt_ms > 0 else None, "peak": hbm_peak,
                           "unit": "GB/s", "frac": (ntt_bytes / (ntt_ms * 1e-3) / 1e9 / hbm_peak) if ntt_ms > 0 else None},
            "phase_ms": phases,
        }
        cpu = None
        if G == 1 and not a.no_cpu_baseline:
            dt, threads, celems = cpu_commit_seconds(w, min(log_n, 16))
            cpu = {"value": celems / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"one full commit of 2^{min(log_n, 16)} rows x {c} cols ({dt:.1f} s), oracle/ built -march=native, OpenMP"}
        h2d = cpr * n * 8 * G
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": G, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64 (Goldilocks field, 32-bit IMAD limbs)", "data": "synthetic",
            "config": {"workload": workload_name(w),
                       "parallelism": "single GPU" if G == 1 else f"coset-sharded x{G}: column-sharded iNTT + " + ("NCCL all-gather of coefficients" if use_nccl else "coefficients stored into every rank's gather buffer by the library's own kernel over NVLink peer memory (flags, no NCCL on the data path)") + " + per-rank cosets/cap subtrees" + ("" if use_nccl else "; e2e: column pipeline (copy -> iNTT -> push of 8/16/rest columns, leaf sponge absorbs chunks as they land)"),
                       "l2": f"inputs rotate over {N_INPUT_SETS} sets ({N_INPUT_SETS * elems * 8 / 1e6:.0f} MB) and each step streams a {8 * N * c / 1e6:.0f} MB LDE, both > 126 MB L2",
                       "elements": "n*c input trace elements per commit"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": (1 << cap) * 32,
                    "ms_per_step": e2e_ms / a.steps, "source": "one pinned contiguous host buffer per rank",
                    # the reference-side shape of the same call: c separately allocated pageable columns, no flatten
                    "pageable_per_column": None if e2e_cols_ms is None else {
                        "value": elems / (e2e_cols_ms / a.steps * 1e-3) / 1e6, "unit": UNIT,
                        "ms_per_step": e2e_cols_ms / a.steps, "entry_point": "vx_commit_from_values_cols",
                        "source": f"{c} separate pageable numpy allocations (plonky2's Vec<PolynomialValues>)"}},
            "shard_parity": shard_parity,
            "prove": prove_block,
            "gpu_launches": launches,
            "wall_ms_per_step": wall_ms / a.steps,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
