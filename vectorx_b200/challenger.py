"""Host-side Fiat-Shamir challenger (plonky2 iop/challenger.rs `Challenger<F, PoseidonHash>`).

north_star keeps the challenger on the host: it absorbs a few hundred elements per proof.  The duplex
sponge needs a scalar Poseidon permutation: the library's host-side vx_challenger_permute."""
from __future__ import annotations

import ctypes

from ._lib import check, load

P = 0xFFFFFFFF00000001


def poseidon_host(state):
    """One scalar permutation on the host (vx_challenger_permute: plain C inside the library, no GPU involved)."""
    buf = (ctypes.c_uint64 * 12)(*[int(x) % P for x in state])
    check(load().vx_challenger_permute(buf), "vx_challenger_permute")
    return [int(x) for x in buf]


def hash_no_pad_host(inputs):
    st = [0] * 12
    for off in range(0, len(inputs), 8):
        chunk = inputs[off:off + 8]
        st[:len(chunk)] = [int(x) % P for x in chunk]
        st = poseidon_host(st)
    return st[:4]


def hash_pad_host(inputs, block=8):
    """`Hasher::hash_pad` (pad10*1 then `hash_no_pad`): push 1, zeros until one short of a multiple of `block`, push 1.
    The only statement of the rule inside the reference tree is the in-tree `Hasher` implementation at
    contracts/lib/succinctx/plonky2x/core/src/backend/wrapper/plonky2_config.rs:174-182, which pads to the number of
    field elements one permutation absorbs (the rate: 8 for Poseidon over Goldilocks, the default here).  plonky2
    releases have also padded to the sponge WIDTH (12); which one v0.2.0 uses cannot be checked in this container, so the
    value is overridable wherever it is consumed (`CircuitData(domain_separator_digest=..., circuit_digest=...)`)."""
    padded = [int(x) % P for x in inputs] + [1]
    while (len(padded) + 1) % block:
        padded.append(0)
    padded.append(1)
    return hash_no_pad_host(padded)


class Challenger:
    def __init__(self):
        self.sponge_state = [0] * 12
        self.input_buffer = []
        self.output_buffer = []

    def _duplexing(self):
        for i, x in enumerate(self.input_buffer):
            self.sponge_state[i] = x
        self.input_buffer = []
        self.sponge_state = poseidon_host(self.sponge_state)
        self.output_buffer = self.sponge_state[:8]

    def observe_element(self, x):
        self.output_buffer = []
        self.input_buffer.append(int(x) % P)
        if len(self.input_buffer) == 8:
            self._duplexing()

    def observe_elements(self, xs):
        for x in xs:
            self.observe_element(x)

    def observe_hash(self, h):
        self.observe_elements(h)

    def observe_cap(self, cap):
        for h in cap:
            self.observe_elements(h)

    def observe_extension_element(self, e):
        self.observe_elements(e)

    def get_challenge(self):
        if self.input_buffer or not self.output_buffer:
            self._duplexing()
        return self.output_buffer.pop()

    def get_n_challenges(self, n):
        return [self.get_challenge() for _ in range(n)]

    def get_extension_challenge(self):
        return [self.get_challenge(), self.get_challenge()]

    def pow_state(self):
        """(state with pending inputs written, position of the witness) for fri_proof_of_work."""
        st = list(self.sponge_state)
        for i, x in enumerate(self.input_buffer):
            st[i] = x
        return st, len(self.input_buffer)
