"""Known-answer tests held by the reference's own test suite for this path, plus a
test-side restatement of the rule construction they need.

The rewrite KATs (reference src/ops/rewrite.zig:244-498) apply
``compileString -> compose -> project(output) -> shortestPath -> printString``; the
minimum-weight output string is unique in each, so they also bind the lazy path's
output string.  Building the rule needs cdrewrite = rmEpsilon((lambda.tau.rho | sigma/1.0)*)
(rewrite.zig:47-95); concat/union/closure/rmEpsilon are restated below from
concat.zig:15-53, union.zig:16-49, closure.zig:23-43 (star), rm-epsilon.zig:15-103.
Grammar construction is OUT OF SCOPE for the product; this file is test infrastructure.
"""
from __future__ import annotations

import math

from common import Spec

INF = math.inf


class M:
    """Tiny mutable FST (arcs in insertion order per state)."""

    def __init__(self):
        self.start = None
        self.finals = []
        self.arcs = []

    def add_state(self):
        self.finals.append(INF); self.arcs.append([]); return len(self.finals) - 1

    def add_states(self, n):
        for _ in range(n): self.add_state()

    def clone(self):
        m = M(); m.start = self.start; m.finals = list(self.finals); m.arcs = [list(a) for a in self.arcs]; return m

    def n(self): return len(self.finals)

    def to_spec(self) -> Spec:
        flat = [(s, il, ol, w, nx) for s, lst in enumerate(self.arcs) for (il, ol, w, nx) in lst]
        return Spec(self.n(), self.start, [None if math.isinf(f) else f for f in self.finals], flat)


def times(a, b):
    return INF if (math.isinf(a) or math.isinf(b)) else a + b


def concat(f1: M, f2: M):   # concat.zig:15-53
    if f2.start is None or f1.start is None:
        return
    off = f1.n()
    f1.add_states(f2.n())
    for s in range(f2.n()):
        if not math.isinf(f2.finals[s]):
            f1.finals[s + off] = f2.finals[s]
        for (il, ol, w, nx) in f2.arcs[s]:
            f1.arcs[s + off].append((il, ol, w, nx + off))
    for s in range(off):
        fw = f1.finals[s]
        if not math.isinf(fw):
            f1.arcs[s].append((0, 0, fw, f2.start + off))
            f1.finals[s] = INF


def union(f1: M, f2: M):    # union.zig:16-49
    if f2.start is None:
        return
    old = f1.start
    off = f1.n()
    f1.add_states(f2.n())
    for s in range(f2.n()):
        if not math.isinf(f2.finals[s]):
            f1.finals[s + off] = f2.finals[s]
        for (il, ol, w, nx) in f2.arcs[s]:
            f1.arcs[s + off].append((il, ol, w, nx + off))
    ns = f1.add_state()
    if old is not None:
        f1.arcs[ns].append((0, 0, 0.0, old))
    f1.arcs[ns].append((0, 0, 0.0, f2.start + off))
    f1.start = ns


def closure_star(f: M):     # closure.zig:23-43
    old = f.start
    if old is None:
        return
    ns = f.add_state()
    f.finals[ns] = 0.0
    f.start = ns
    f.arcs[ns].append((0, 0, 0.0, old))
    for s in range(ns):
        if not math.isinf(f.finals[s]):
            f.arcs[s].append((0, 0, f.finals[s], old))


def rm_epsilon(f: M) -> M:  # rm-epsilon.zig:15-103
    r = M()
    if f.n() == 0 or f.start is None:
        return r
    r.add_states(f.n())
    r.start = f.start
    for s in range(f.n()):
        cs, cw = [], []
        visited = [False] * f.n()
        stack = [(s, 0.0)]
        visited[s] = True
        while stack:
            cur, w = stack.pop()
            cs.append(cur); cw.append(w)
            for (il, ol, aw, nx) in f.arcs[cur]:
                if il == 0 and ol == 0 and not visited[nx]:
                    visited[nx] = True
                    stack.append((nx, times(w, aw)))
        fw = f.finals[s]
        for st, w in zip(cs, cw):
            if st == s:
                continue
            if not math.isinf(f.finals[st]):
                fw = min(fw, times(w, f.finals[st]))
        if not math.isinf(fw):
            r.finals[s] = fw
        for st, w in zip(cs, cw):
            for (il, ol, aw, nx) in f.arcs[st]:
                if il != 0 or ol != 0:
                    r.arcs[s].append((il, ol, times(w, aw), nx))
    return r


def trivial_eps(f: M):      # rewrite.zig:137-140
    if f.start is None:
        return True
    return f.n() == 1 and not math.isinf(f.finals[f.start]) and not f.arcs[f.start]


def cdrewrite(tau: M, lam: M, rho: M, sigma_labels) -> M:   # rewrite.zig:47-135
    lt, rt = trivial_eps(lam), trivial_eps(rho)
    if lt and rt:
        ctx = tau.clone()
    elif lt:
        ctx = tau.clone(); concat(ctx, rho)
    elif rt:
        ctx = lam.clone(); concat(ctx, tau)
    else:
        ctx = lam.clone(); concat(ctx, tau); concat(ctx, rho)
    one = M(); s0 = one.add_state(); s1 = one.add_state(); one.start = s0; one.finals[s1] = 0.0
    for l in sorted(sigma_labels):
        one.arcs[s0].append((l, l, 1.0, s1))   # IDENTITY_PENALTY rewrite.zig:21
    union(ctx, one)
    closure_star(ctx)
    return rm_epsilon(ctx)


def lab(ch: str) -> int:
    return ord(ch) + 1


def linear(pairs) -> M:
    """Chain of (in_char, out_char) arcs, unit weight, last state final."""
    m = M(); m.add_states(len(pairs) + 1); m.start = 0; m.finals[len(pairs)] = 0.0
    for i, (a, b) in enumerate(pairs):
        m.arcs[i].append((lab(a), lab(b), 0.0, i + 1))
    return m


def eps_fst() -> M:
    m = M(); m.add_state(); m.start = 0; m.finals[0] = 0.0; return m


SIGMA = [lab(chr(c)) for c in range(ord("a"), ord("z") + 1)]   # rewrite.zig:220-233


def branching_lambda() -> M:   # rewrite.zig:457-465
    m = M(); m.add_states(2); m.start = 0; m.finals[1] = 0.0
    m.arcs[0].append((lab("c"), lab("c"), 0.0, 1)); m.arcs[0].append((lab("x"), lab("x"), 0.0, 1))
    return m


def rewrite_kats():
    """[(rule name, rule M, [(input, expected output)])] — rewrite.zig:244-498."""
    a_b = linear([("a", "b")])
    ab_xy = linear([("a", "x"), ("b", "y")])
    c, d = linear([("c", "c")]), linear([("d", "d")])
    return [
        ("a->b/_", cdrewrite(a_b, eps_fst(), eps_fst(), SIGMA), [(b"a", b"b"), (b"hello", b"hello")]),
        ("ab->xy/_", cdrewrite(ab_xy, eps_fst(), eps_fst(), SIGMA), [(b"ab", b"xy"), (b"ac", b"ac"), (b"cab", b"cxy"), (b"aab", b"axy")]),
        ("a->b/c_d", cdrewrite(a_b, c, d, SIGMA), [(b"cad", b"cbd"), (b"cab", b"cab"), (b"xad", b"xad")]),
        ("ab->xy/c_d", cdrewrite(ab_xy, c, d, SIGMA), [(b"cabd", b"cxyd"), (b"cacd", b"cacd")]),
        ("a->b/(c|x)_d", cdrewrite(a_b, branching_lambda(), d, SIGMA), [(b"cad", b"cbd"), (b"xad", b"xbd"), (b"yad", b"yad")]),
    ]
