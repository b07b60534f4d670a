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


# ── config 4: synthetic WeText-style tagger (not defined by the reference) ──
_ALNUM = b"abcdefghijklmnopqrstuvwxyz0123456789"


def wetext_arrays(K: int = 110000, seed: int = 20261018):
    """Generator of BASELINE config 4 as plain arrays: (n_states, src, ilabel, olabel, weight, nextstate, sources);
    start state 0, the only final state is 0 (weight 0).  `sources` are the dictionary keys.

    state 0: start, final 0, 256 identity self-loops (c+1 : c+1 / 1.0 = rewrite.zig IDENTITY_PENALTY);
    an open insertion chain 0 -e:'t'-> -e:'a'-> -e:'g'-> -e:'{'-> trie root (weights 0);
    K entries src -> dst (src U[2,8], dst U[0,10] chars over [a-z0-9], weight U{0..8}/8, every fifth entry re-uses
    an earlier src); src goes into a shared input trie (ch+1 : 0 / 0); from the word-end node an epsilon-input
    chain emits dst (first arc carries the weight), then (0:0) to a shared close state (carrying the weight if
    dst is empty), close -e:'}'-> -e:e-> state 0.
    """
    rng = np.random.default_rng(seed)
    src_a, il_a, ol_a, w_a, nx_a = [], [], [], [], []

    def arc(s, il, ol, w, n):
        src_a.append(s); il_a.append(il); ol_a.append(ol); w_a.append(w); nx_a.append(n)

    n_states = 1
    for c in range(256):
        arc(0, c + 1, c + 1, 1.0, 0)
    prev = 0
    for ch in b"tag{":
        arc(prev, 0, ch + 1, 0.0, n_states); prev = n_states; n_states += 1
    root = prev
    close = n_states; n_states += 1
    after = n_states; n_states += 1
    arc(close, 0, ord("}") + 1, 0.0, after)
    arc(after, 0, 0, 0.0, 0)
    trie = {}          # (node, byte) -> node
    sources = []
    lens_s = rng.integers(2, 9, K); lens_d = rng.integers(0, 11, K); ws = rng.integers(0, 9, K) / 8.0
    for k in range(K):
        if k % 5 == 4 and sources:
            s = sources[int(rng.integers(0, len(sources)))]
        else:
            s = bytes(_ALNUM[i] for i in rng.integers(0, 36, int(lens_s[k])))
        sources.append(s)
        d = bytes(_ALNUM[i] for i in rng.integers(0, 36, int(lens_d[k])))
        node = root
        for ch in s:
            nxt = trie.get((node, ch))
            if nxt is None:
                nxt = n_states; n_states += 1
                trie[(node, ch)] = nxt
                arc(node, ch + 1, 0, 0.0, nxt)
            node = nxt
        w = float(ws[k])
        cur = node
        for j, ch in enumerate(d):
            arc(cur, 0, ch + 1, w if j == 0 else 0.0, n_states); cur = n_states; n_states += 1
        arc(cur, 0, 0, w if len(d) == 0 else 0.0, close)
    return (n_states, np.array(src_a, np.uint32), np.array(il_a, np.uint32), np.array(ol_a, np.uint32), np.array(w_a, np.float64),
            np.array(nx_a, np.uint32), sources)


def wetext_style(K: int = 110000, seed: int = 20261018):
    """The config-4 transducer built through the product C ABI: (MutableFst, sources)."""
    n_states, src, il, ol, w, nxt, sources = wetext_arrays(K, seed)
    m = MutableFst()
    m.add_states(n_states)
    m.set_start(0)
    m.set_finals(np.array([0], np.uint32), np.zeros(1))
    rc = m.add_arcs(src, il, ol, w, nxt)
    assert rc == 0, rc
    return m, sources


def wetext_strings(sources, n: int, seed: int = 1, lo: int = 11, hi: int = 251):
    """n input strings: length L ~ U[lo,hi]; 70 % concatenated dictionary sources until length >= L, 30 % uniform
    printable bytes."""
    rng = np.random.default_rng(seed)
    out = []
    Ls = rng.integers(lo, hi + 1, n); kinds = rng.random(n) < 0.7
    for i in range(n):
        L = int(Ls[i])
        if kinds[i]:
            parts, tot = [], 0
            while tot < L:
                s = sources[int(rng.integers(0, len(sources)))]
                parts.append(s); tot += len(s)
            out.append(b"".join(parts))
        else:
            out.append(bytes(rng.integers(32, 127, L).astype(np.uint8)))
    return out


def wetext_packed(sources, n: int, seed: int = 1, lo: int = 11, hi: int = 251, block: int = 1 << 17):
    """The same distribution as wetext_strings, vectorised for large batches: (uint8 data, uint64 offsets[n + 1]).
    Dictionary strings are grown in rounds (every string still shorter than its target length L appends one random
    dictionary source per round), random strings are uniform printable bytes; generated block by block so that the
    index arrays stay small.  Deterministic in (sources, n, seed, lo, hi, block)."""
    rng = np.random.default_rng(seed)
    src_len = np.fromiter((len(s) for s in sources), np.int64, len(sources))
    src_off = np.zeros(len(sources) + 1, np.int64); np.cumsum(src_len, out=src_off[1:])
    src_bytes = np.frombuffer(b"".join(sources), np.uint8)
    datas, lens_all = [], []
    for b0 in range(0, n, block):
        m = min(block, n - b0)
        Ls = rng.integers(lo, hi + 1, m).astype(np.int64)
        is_dict = rng.random(m) < 0.7
        tot = np.zeros(m, np.int64)
        tok_str, tok_src, tok_pos = [], [], []
        active = np.flatnonzero(is_dict)
        while active.size:
            pick = rng.integers(0, len(sources), active.size)
            tok_str.append(active); tok_src.append(pick); tok_pos.append(tot[active].copy())
            tot[active] += src_len[pick]
            active = active[tot[active] < Ls[active]]
        lens = np.where(is_dict, tot, Ls)
        off = np.zeros(m + 1, np.int64); np.cumsum(lens, out=off[1:])
        data = np.empty(int(off[-1]), np.uint8)
        # uniform printable bytes for the other strings
        rid = np.flatnonzero(~is_dict)
        if rid.size:
            rl = lens[rid]
            start = np.repeat(off[rid], rl)
            within = np.arange(int(rl.sum()), dtype=np.int64) - np.repeat(np.cumsum(rl) - rl, rl)
            data[start + within] = rng.integers(32, 127, int(rl.sum())).astype(np.uint8)
        if tok_str:
            ts, tsrc, tpos = np.concatenate(tok_str), np.concatenate(tok_src), np.concatenate(tok_pos)
            sl = src_len[tsrc]
            within = np.arange(int(sl.sum()), dtype=np.int64) - np.repeat(np.cumsum(sl) - sl, sl)
            data[np.repeat(off[ts] + tpos, sl) + within] = src_bytes[np.repeat(src_off[tsrc], sl) + within]
        datas.append(data); lens_all.append(lens)
    lens = np.concatenate(lens_all) if lens_all else np.zeros(0, np.int64)
    offsets = np.zeros(n + 1, np.uint64); np.cumsum(lens.astype(np.uint64), out=offsets[1:])
    return (np.concatenate(datas) if datas else np.zeros(0, np.uint8)), offsets
