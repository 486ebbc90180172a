"""LocalProver (mirror of contracts/lib/succinctx/plonky2x/core/src/backend/prover/local.rs): the proof-level fan-out
of batch_prove over devices.  CPU tests drive the scheduler with a stand-in engine; the GPU test proves a batch on real
contexts (two contexts on one GPU when only one is visible) and checks every proof against the oracle."""
import threading
import time

import pytest

import vectorx_b200 as vx
from vectorx_b200.local_prover import CircuitSpec, LocalProver


class FakeCtx:
    made = []

    def __init__(self, device):
        self.device = device
        FakeCtx.made.append(device)

    def close(self):
        pass


class FakeSpec:
    binds = []

    def bind(self, ctx):
        FakeSpec.binds.append(ctx.device)
        return ("circuit-on", ctx.device)


def fake_prove(circ, wires, public_inputs):
    time.sleep(0.01 * (1 + wires % 3))
    return {"device": circ[1], "wires": wires, "pi": public_inputs}


def test_batch_prove_keeps_input_order_and_uses_every_device():
    FakeCtx.made.clear(); FakeSpec.binds.clear()
    lp = LocalProver(devices=[0, 1, 2, 3], make_ctx=FakeCtx, prove_fn=fake_prove)
    spec = FakeSpec()
    inputs = [(i, [i, i + 1]) for i in range(32)]             # the 32 map proofs of a header_range job
    out = lp.batch_prove(spec, inputs)
    assert [o["wires"] for o in out] == list(range(32)) and [o["pi"] for o in out] == [[i, i + 1] for i in range(32)]
    assert sorted(FakeCtx.made) == [0, 1, 2, 3] and sorted(FakeSpec.binds) == [0, 1, 2, 3]      # one replica per device
    assert set(lp.last_assignment) == {0, 1, 2, 3}
    assert all(out[i]["device"] == lp.devices[lp.last_assignment[i]] for i in range(32))
    out2 = lp.batch_prove(spec, inputs[:5])                   # replicas are reused, not rebuilt
    assert len(FakeSpec.binds) == 4 and [o["wires"] for o in out2] == [0, 1, 2, 3, 4]
    assert lp.batch_prove(spec, []) == []
    single = lp.prove(spec, (7, [1]))
    assert single["wires"] == 7 and single["device"] == 0
    lp.close()


def test_batch_prove_fewer_inputs_than_devices_and_errors():
    FakeCtx.made.clear()
    lp = LocalProver(devices=[0, 1, 2, 3, 4, 5, 6, 7], make_ctx=FakeCtx, prove_fn=fake_prove)
    out = lp.batch_prove(FakeSpec(), [(1, []), (2, [])])
    assert [o["wires"] for o in out] == [1, 2] and len(FakeCtx.made) == 2

    def bad(circ, wires, pi):
        if wires == 3:
            raise ValueError("witness does not satisfy the circuit")
        return fake_prove(circ, wires, pi)
    lp2 = LocalProver(devices=[0, 1], make_ctx=FakeCtx, prove_fn=bad)
    with pytest.raises(RuntimeError, match="input 3 failed.*does not satisfy"):
        lp2.batch_prove(FakeSpec(), [(i, []) for i in range(8)])
    with pytest.raises(ValueError):
        LocalProver(devices=[], make_ctx=FakeCtx)


def test_workers_run_concurrently():
    gate, seen = threading.Barrier(3, timeout=5), []

    def rendezvous(circ, wires, pi):                         # completes only if three proofs are in flight at once
        gate.wait()
        seen.append(circ[1])
        return wires
    lp = LocalProver(devices=[0, 1, 2], make_ctx=FakeCtx, prove_fn=rendezvous)
    assert lp.batch_prove(FakeSpec(), [(i, []) for i in range(3)]) == [0, 1, 2]
    assert sorted(seen) == [0, 1, 2]


def test_workers_per_device_share_one_context_and_replica():
    """workers_per_device = 2: four proofs in flight on two devices, ONE context and ONE bound circuit per device (the
    context's lanes carry the concurrent calls)."""
    FakeCtx.made.clear(); FakeSpec.binds.clear()
    gate, seen = threading.Barrier(4, timeout=5), []

    def rendezvous(circ, wires, pi):                         # completes only if four proofs are in flight at once
        gate.wait()
        seen.append(circ[1])
        return wires
    lp = LocalProver(devices=[0, 1], make_ctx=FakeCtx, prove_fn=rendezvous, workers_per_device=2)
    assert lp.batch_prove(FakeSpec(), [(i, []) for i in range(4)]) == [0, 1, 2, 3]
    assert sorted(seen) == [0, 0, 1, 1]
    assert sorted(FakeCtx.made) == [0, 1] and sorted(FakeSpec.binds) == [0, 1]
    with pytest.raises(ValueError):
        LocalProver(devices=[0], make_ctx=FakeCtx, workers_per_device=0)


def test_no_gpu_means_no_prover():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert vx.device_count() == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        LocalProver()


@pytest.mark.gpu
def test_batch_prove_on_devices_matches_oracle():
    from oracle import plonk, synth
    from test_gpu_prover import normalise, to_oracle_proof
    n_dev = vx.device_count()
    assert n_dev >= 1
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]    # one GPU: two contexts (streams) on it
    circ, wires, pis = synth.build(6, seed=21)
    spec = CircuitSpec(circ.d, [g.id() for g in circ.gates], circ.selector_index, circ.groups, circ.constants, circ.sigmas)
    circ_b, wires_b, pis_b = synth.build(6, seed=22)           # a second circuit of the same shape: its own replicas
    spec_b = CircuitSpec(circ_b.d, [g.id() for g in circ_b.gates], circ_b.selector_index, circ_b.groups,
                         circ_b.constants, circ_b.sigmas)
    inputs = [(wires, pis)] * 6                              # proving is deterministic: six identical proofs expected
    lp = LocalProver(devices=devices)
    proofs = lp.batch_prove(spec, inputs)
    assert len(proofs) == len(inputs) and set(lp.last_assignment) <= set(range(len(devices)))
    want = normalise(plonk.prove(circ, wires, pis))
    for pr in proofs:
        assert normalise(pr) == want
        assert plonk.verify(circ, to_oracle_proof(pr))
    proofs_b = lp.batch_prove(spec_b, [(wires_b, pis_b)] * 3)
    want_b = normalise(plonk.prove(circ_b, wires_b, pis_b))
    assert all(normalise(pr) == want_b for pr in proofs_b)
    assert normalise(lp.prove(spec, (wires, pis))) == want
    lp.close()
