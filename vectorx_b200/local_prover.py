"""Host-side mirror of the reference's `LocalProver` (contracts/lib/succinctx/plonky2x/core/src/backend/prover/local.rs):
`prove(circuit, input)` and `batch_prove(circuit, inputs)`.

The reference's `batch_prove` is a sequential `for` over independent inputs (local.rs:44-48) -- the 32 map proofs and
the 16/8/4/2/1 reduce proofs of a header_range job (frontend/mapreduce/generator.rs:109-146).  Here the independent
proofs fan out over the visible B200s: one worker thread, one `vx_ctx` and one resident replica of the
circuit's prover data (constants/sigmas commitment, sigma columns, gate program) per device; workers pull the next
input from a shared queue, so a slow proof never idles the other GPUs.  There is no data-path collective: proofs are
independent ("replicas" axis of DESIGN.md section 4).  Every C-ABI call releases the GIL, so the workers overlap.
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, Sequence

from ._lib import Context, device_list
from .prover import CircuitData, prove


class CircuitSpec:
    """What `CircuitBuild` holds that this path needs, not yet bound to a device: the arguments of `CircuitData`."""

    def __init__(self, degree_bits, gate_ids, selector_index, groups, constants, sigmas, **config):
        self.args = (degree_bits, list(gate_ids), list(selector_index), [tuple(g) for g in groups], constants, sigmas)
        self.config = dict(config)

    def bind(self, ctx: Context) -> CircuitData:
        return CircuitData(*self.args, ctx=ctx, **self.config)


class _Replica:
    """One device's context and the circuits already resident on it."""

    def __init__(self, device: int, make_ctx: Callable[[int], object]):
        self.device = device
        self.ctx = make_ctx(device)
        self.bound: dict[int, object] = {}

    def circuit(self, spec):
        key = id(spec)
        if key not in self.bound:
            self.bound[key] = spec.bind(self.ctx)
        return self.bound[key]


class LocalProver:
    """`LocalProver::new()`; `devices` defaults to every sm_100-class GPU the library sees.  A device may be listed more
    than once: each entry gets its own context (stream), which keeps one GPU busy across the host-side gaps of a proof."""

    def __init__(self, devices: Sequence[int] | None = None, *, make_ctx: Callable[[int], object] = Context,
                 prove_fn: Callable = prove, workers_per_device: int = 1):
        if devices is None:
            devices = device_list()           # the usable CUDA ordinals (a box may mix GPU generations)
            if len(devices) == 0:
                raise RuntimeError("LocalProver: no sm_100-class GPU is visible (there is no CPU fallback)")
        if len(devices) == 0:
            raise ValueError("LocalProver: empty device list")
        # workers_per_device > 1: that many worker threads share ONE context (and one resident circuit replica) per listed
        # device -- the context's lanes run their calls side by side, which fills the host-side gaps of a proof
        if workers_per_device < 1:
            raise ValueError("LocalProver: workers_per_device must be >= 1")
        self.workers_per_device = int(workers_per_device)
        self.devices = list(devices)
        self._make_ctx, self._prove_fn = make_ctx, prove_fn
        self._replicas: list[_Replica | None] = [None] * len(self.devices)
        self.last_assignment: list[int] = []            # worker slot that proved each input of the last batch

    def _replica(self, slot: int) -> _Replica:
        if self._replicas[slot] is None:
            self._replicas[slot] = _Replica(self.devices[slot], self._make_ctx)
        return self._replicas[slot]

    def prove(self, circuit: CircuitSpec, input):
        """`input` = (wires, public_inputs): the witness matrix the (host-side) generators produced and the circuit's
        public inputs.  Returns the proof."""
        wires, public_inputs = input
        return self._prove_fn(self._replica(0).circuit(circuit), wires, public_inputs)

    def batch_prove(self, circuit: CircuitSpec, inputs: Sequence) -> list:
        """Proofs of all inputs, in input order.  The first failure is re-raised after the workers have drained."""
        inputs = list(inputs)
        results: list = [None] * len(inputs)
        self.last_assignment = [-1] * len(inputs)
        if not inputs:
            return results
        jobs: queue.SimpleQueue = queue.SimpleQueue()
        for i in range(len(inputs)):
            jobs.put(i)
        errors: list[tuple[int, BaseException]] = []
        stop = threading.Event()

        bind_lock = threading.Lock()

        def worker(slot: int):
            try:
                with bind_lock:                          # workers of one device share its replica: bind it once
                    circ = self._replica(slot).circuit(circuit)
            except BaseException as e:                   # noqa: BLE001 -- reported to the caller below
                errors.append((-1, e))
                stop.set()
                return
            while not stop.is_set():
                try:
                    i = jobs.get_nowait()
                except queue.Empty:
                    return
                try:
                    wires, public_inputs = inputs[i]
                    results[i] = self._prove_fn(circ, wires, public_inputs)
                    self.last_assignment[i] = slot
                except BaseException as e:               # noqa: BLE001
                    errors.append((i, e))
                    stop.set()
                    return

        slots = [s for s in range(len(self.devices)) for _ in range(self.workers_per_device)][:len(inputs)]
        threads = [threading.Thread(target=worker, args=(s,), name=f"vx-prover-{self.devices[s]}") for s in slots]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            i, e = min(errors, key=lambda p: p[0])
            raise RuntimeError(f"batch_prove: input {i} failed on a worker: {e}") from e
        return results

    def close(self):
        for r in self._replicas:
            if r is not None:
                for circ in r.bound.values():            # device data first, then the context it lives on
                    release = getattr(circ, "close", None)
                    if release:
                        release()
                r.bound.clear()
                close = getattr(r.ctx, "close", None)
                if close:
                    close()
        self._replicas = [None] * len(self.devices)
