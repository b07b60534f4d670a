"""Synthetic workloads of the reference's bench scenarios, built through the C ABI.

Definitions follow reference bench/optimize-bench.zig (generators :164-277,
inputs :279-328); they are restated here with numpy so that bench.py does not
need the oracle to create its inputs.
"""
from __future__ import annotations

import numpy as np

from . import MutableFst


def _build(num_states, start, final_states, src, il, ol, w, nxt):
    m = MutableFst()
    m.add_states(int(num_states))
    m.set_start(int(start))
    fs = np.asarray(final_states, np.uint32)
    m.set_finals(fs, np.zeros(len(fs)))
    rc = m.add_arcs(src, il, ol, w, nxt)
    assert rc == 0, rc
    return m


def plain_transducer(T: int, B: int) -> MutableFst:
    """transducer_frozen (optimize-bench.zig:290-305): deterministic on input."""
    i = np.repeat(np.arange(T, dtype=np.int64), B)
    b = np.tile(np.arange(B, dtype=np.int64), T)
    return _build(T, 0, np.arange(T), i, (b % 255) + 1, ((i + b) % 255) + 1, b.astype(np.float64), (i + b + 1) % T)


def epsilon_dense_transducer(T: int, B: int) -> MutableFst:
    """buildEpsilonDenseTransducer (optimize-bench.zig:219-248)."""
    i = np.repeat(np.arange(T, dtype=np.int64), B + 1)
    k = np.tile(np.arange(B + 1, dtype=np.int64), T)     # k == 0: the epsilon arc, k-1 = b
    b = np.maximum(k - 1, 0)
    eps = k == 0
    il = np.where(eps, 0, 1)
    ol = np.where(eps, 0, ((i + b) % 255) + 1)
    w = np.where(eps, 0.0, b.astype(np.float64))
    nxt = np.where(eps, i + 1, np.minimum(i + (b % 4) + 1, T))
    return _build(T + 1, 0, np.arange(T + 1), i, il, ol, w, nxt)


def ambiguous_chain_transducer(T: int, B: int) -> MutableFst:
    """buildAmbiguousChainTransducer (optimize-bench.zig:250-277)."""
    fan = max(1, min(B, 4))
    i = np.repeat(np.arange(T + 1, dtype=np.int64), fan + 1)
    k = np.tile(np.arange(fan + 1, dtype=np.int64), T + 1)   # k == 0: the stay arc
    b = np.maximum(k - 1, 0)
    stay = k == 0
    il = np.ones_like(i)
    ol = np.where(stay, 1, ((i + b) % 255) + 1)
    w = np.where(stay, 0.0, b.astype(np.float64))
    nxt = np.where(stay, i, np.minimum(i + b + 1, T))
    return _build(T + 1, 0, np.arange(T + 1), i, il, ol, w, nxt)


TRANSDUCERS = {"plain": plain_transducer, "epsilon_dense": epsilon_dense_transducer, "ambiguous": ambiguous_chain_transducer}


def input_string(workload: str, length: int, branches: int) -> bytes:
    """Bytes whose compiled acceptor (label = byte + 1) equals the scenario's left operand:
    acceptor_repeat = label 1 x len (:182-196, :284); acceptor_branch = label (i % B) + 1 (:164-180, :282)."""
    if workload == "plain":
        return bytes(i % max(1, branches) for i in range(length))
    return bytes(length)
