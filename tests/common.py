"""Shared helpers for the parity tests: seeded generators and dual construction
(the same transducer built through the product C ABI and through the oracle)."""
from __future__ import annotations

import random

import numpy as np


class Spec:
    """A transducer / acceptor description: arcs in insertion order."""

    def __init__(self, num_states, start, finals, arcs):
        self.num_states, self.start, self.finals, self.arcs = num_states, start, list(finals), list(arcs)

    def to_product(self, L):
        m = L.MutableFst()
        m.add_states(self.num_states)
        if self.start is not None:
            m.set_start(self.start)
        for s, w in enumerate(self.finals):
            if w is not None:
                m.set_final(s, w)
        for (src, il, ol, w, nxt) in self.arcs:
            assert m.add_arc(src, il, ol, w, nxt) == 0
        return m

    def to_oracle(self, O):
        m = O.Mutable()
        m.add_states(self.num_states)
        if self.start is not None:
            m.set_start(self.start)
        for s, w in enumerate(self.finals):
            if w is not None:
                m.set_final(s, w)
        for (src, il, ol, w, nxt) in self.arcs:
            m.add_arc(src, il, ol, w, nxt)
        return m


def random_rhs(rng: random.Random, max_states=8, nlab=None, eps_p=None, wmax=None, neg=False, real=False) -> Spec:
    """SURVEY App. C generator: tie-heavy small transducers with epsilons."""
    n = rng.randint(2, max_states)
    nlab = nlab or rng.randint(1, 3)
    eps_p = rng.choice([0.0, 0.2, 0.4]) if eps_p is None else eps_p
    wmax = rng.choice([0, 1, 2, 3]) if wmax is None else wmax
    narcs = rng.randint(n, 5 * n)
    arcs = []
    for _ in range(narcs):
        src = rng.randrange(n)
        il = 0 if rng.random() < eps_p else rng.randint(1, nlab)
        ol = rng.randint(0, nlab)
        if real:
            w = rng.choice([0.0, 0.125, 0.25, 0.5, 1.0, 1.1, 2.3, 0.1, 0.2, 0.3])
        else:
            w = float(rng.randint(-wmax if neg else 0, wmax))
        arcs.append((src, il, ol, w, rng.randrange(n)))
    finals = [float(rng.randint(0, wmax)) if rng.random() < 0.4 else None for _ in range(n)]
    return Spec(n, 0, finals, arcs)


def random_string(rng: random.Random, nlab=3, max_len=6) -> bytes:
    return bytes(rng.randint(0, nlab - 1) for _ in range(rng.randint(0, max_len)))


def random_lhs(rng: random.Random, max_states=6, nlab=3, eps_p=0.25, wmax=2, neg=False) -> Spec:
    """General (non-linear) left operand with output epsilons."""
    n = rng.randint(1, max_states)
    narcs = rng.randint(0, 3 * n)
    arcs = []
    for _ in range(narcs):
        src = rng.randrange(n)
        il = rng.randint(0, nlab)
        ol = 0 if rng.random() < eps_p else rng.randint(1, nlab)
        w = float(rng.randint(-wmax if neg else 0, wmax))
        arcs.append((src, il, ol, w, rng.randrange(n)))
    finals = [float(rng.randint(0, wmax)) if rng.random() < 0.5 else None for _ in range(n)]
    return Spec(n, 0, finals, arcs)


def frozen_pair(L, O, spec: Spec):
    """Freeze through the product ABI and load the SAME image into the oracle."""
    import os
    import tempfile
    fm = spec.to_product(L)
    f = fm.freeze()
    with tempfile.NamedTemporaryFile(suffix=".fst", delete=False) as t:
        p = t.name
    try:
        assert f.save(p) == 0
        img = open(p, "rb").read()
    finally:
        os.unlink(p)
    return f, O.Frozen.from_bytes(img), img


def gen_image(O, kind, T, B) -> bytes:
    return O.Frozen.generate(kind, T, B).to_bytes()


def assert_batch_matches_oracle(L, O, fprod, forc, strings, res=None, check_out=True):
    """Compare a product batch result with the oracle string by string (bit-exact)."""
    if res is None:
        data, offsets = L.pack_strings(strings)
        res = L.compose_frozen_shortest_path_batch(fprod, data, offsets)
    assert len(res.status) == len(strings)
    for i, s in enumerate(strings):
        p = O.csp_bytes(forc, s)
        if p.status == O.STATUS_BACKTRACK_CYCLE:
            assert res.status[i] == L.CYCLE, (i, s, res.status[i])
            continue
        if p.status == O.STATUS_EMPTY:
            assert res.status[i] == L.NO_PATH, (i, s, res.status[i])
            continue
        not_bytes = bool((p.olabels > 256).any())   # no byte form of the output tape (string.zig:64-97 -> null)
        assert res.status[i] == (L.NOT_BYTES if not_bytes else L.PATH), (i, s, res.status[i], p.status)
        il, ol, w = res.path(i)
        assert np.array_equal(il, p.ilabels), (i, s, il, p.ilabels)
        assert np.array_equal(ol, p.olabels), (i, s, ol, p.olabels)
        assert np.array_equal(w.view(np.uint64), p.weights.view(np.uint64)), (i, s, w, p.weights)
        a, b = np.float64(res.final_weights[i]), np.float64(p.final_weight)
        assert a.view(np.uint64) == b.view(np.uint64), (i, s, a, b)
        if check_out:
            assert res.output(i) == (b"" if not_bytes else p.output_bytes()), (i, s)
    return res


def assert_batch_matches_eager_oracle(L, O, fprod, forc, strings, res=None):
    """Eager semantics (BASELINE config 5): the product batch (configured with semantics=EAGER) against the oracle's
    compose() followed by shortestPath() (compose.zig:29-198, shortest-path.zig:18-139), bit-exact."""
    if res is None:
        data, offsets = L.pack_strings(strings)
        res = L.compose_frozen_shortest_path_batch(fprod, data, offsets)
    for i, s in enumerate(strings):
        p, lat_states, lat_arcs = O.eager_mutable(O.Mutable.compile_string(s), forc, 1)
        if p.status == O.STATUS_BACKTRACK_CYCLE:
            assert res.status[i] == L.CYCLE, (i, s, res.status[i])
            continue
        if p.status == O.STATUS_EMPTY:
            assert res.status[i] == L.NO_PATH, (i, s, res.status[i])
            continue
        assert res.status[i] == L.PATH, (i, s, res.status[i], p.status)
        il, ol, w = res.path(i)
        assert np.array_equal(il, p.ilabels), (i, s, il, p.ilabels)
        assert np.array_equal(ol, p.olabels), (i, s, ol, p.olabels)
        assert np.array_equal(w.view(np.uint64), p.weights.view(np.uint64)), (i, s, w, p.weights)
        a, b = np.float64(res.final_weights[i]), np.float64(p.final_weight)
        assert a.view(np.uint64) == b.view(np.uint64), (i, s, a, b)
        assert res.n_tuples[i] == lat_states, (i, s, res.n_tuples[i], lat_states)   # the whole lattice was numbered
    return res
