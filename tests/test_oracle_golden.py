"""CPU: pin the oracle against every golden vector the reference holds for this path
and against the committed fixtures (tests/golden)."""
import json
import os

import numpy as np
import pytest

import refkat
from common import Spec

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_unit_vector_lazy_equals_eager(O):
    # compose-shortest-path.zig:424-471: "123" o ("123" -> "abc"), mutable and frozen rhs agree
    rhs = O.Mutable.compile_string_transducer(b"123", b"abc").freeze()
    lhs = O.Mutable.compile_string(b"123")
    lazy = O.csp_mutable(lhs, rhs, 1)
    eager, ls, la = O.eager_mutable(lhs, rhs, 1)
    for p in (lazy, eager):
        assert p.status == O.STATUS_OK
        assert list(p.ilabels) == [50, 51, 52] and list(p.olabels) == [98, 99, 100] and list(p.weights) == [0, 0, 0]
        assert p.final_weight == 0.0
    assert lazy.final_tuple == (3, 3, 0) and (ls, la) == (4, 3)
    assert O.csp_mutable(lhs, rhs, 0).status == O.STATUS_EMPTY
    assert O.csp_mutable(lhs, rhs, 2).status == O.STATUS_UNSUPPORTED_N    # :33, diff-test.zig:278-287


def test_shortest_path_unit_vectors(O):
    # shortest-path.zig:143-176: best of two paths has weight 3.0, first arc weight 1.0
    spec = Spec(4, 0, [None, None, 0.0, None], [(0, 1, 1, 1.0, 1), (1, 2, 2, 2.0, 2), (0, 3, 3, 5.0, 3), (3, 4, 4, 1.0, 2)])
    ident = Spec(1, 0, [0.0], [(0, l, l, 0.0, 0) for l in (1, 2, 3, 4)])   # sigma* on the labels: compose is the identity
    p, _, _ = O.eager_mutable(spec.to_oracle(O), ident.to_oracle(O).freeze(), 1)
    assert p.status == O.STATUS_OK and len(p.ilabels) == 2 and p.weights[0] == 1.0 and p.total == 3.0
    # :191-216: cheaper of two parallel arcs to the same state
    spec = Spec(2, 0, [None, 0.0], [(0, 10, 100, 3.0, 1), (0, 11, 101, 1.0, 1)])
    ident = Spec(1, 0, [0.0], [(0, 100, 100, 0.0, 0), (0, 101, 101, 0.0, 0)])
    for p in (O.eager_mutable(spec.to_oracle(O), ident.to_oracle(O).freeze(), 1)[0], O.csp_mutable(spec.to_oracle(O), ident.to_oracle(O).freeze(), 1)):
        assert p.status == O.STATUS_OK and list(p.ilabels) == [11] and list(p.olabels) == [101] and list(p.weights) == [1.0]


def test_compose_frozen_lookup_vector(O):
    # compose.zig:311-354 — duplicate ilabels in the rhs; lattice has 3 states / 3 arcs
    lhs = Spec(3, 0, [None, None, 0.0], [(0, 1, 10, 0.0, 1), (1, 2, 20, 0.0, 2)])
    rhs = Spec(3, 0, [None, None, 0.0], [(0, 5, 50, 0.0, 1), (0, 10, 100, 0.0, 1), (0, 10, 101, 2.0, 1), (0, 15, 150, 0.0, 1),
                                         (1, 20, 200, 0.0, 2), (1, 21, 201, 0.0, 2)])
    p, ls, la = O.eager_mutable(lhs.to_oracle(O), rhs.to_oracle(O).freeze(), 1)
    assert (ls, la) == (3, 3) and list(p.olabels) == [100, 200] and p.total == 0.0
    lazy = O.csp_mutable(lhs.to_oracle(O), rhs.to_oracle(O).freeze(), 1)
    assert list(lazy.olabels) == [100, 200] and list(lazy.ilabels) == [1, 2]


def test_rewrite_kats(O):
    # rewrite.zig:244-498 through both the eager pair (what the reference's test runs) and the lazy path
    n = 0
    for name, rule, cases in refkat.rewrite_kats():
        f = rule.to_spec().to_oracle(O).freeze()
        for inp, want in cases:
            lhs = O.Mutable.compile_string(inp)
            eager, _, _ = O.eager_mutable(lhs, f, 1)
            lazy = O.csp_mutable(lhs, f, 1)
            assert eager.output_bytes() == want, (name, inp, eager.output_bytes())
            assert lazy.output_bytes() == want, (name, inp, lazy.output_bytes())
            assert lazy.total == eager.total
            n += 1
    assert n == 14


def test_optimize_kat(O):
    # optimize.zig:164-201: "a" through (a:b) gives "b"
    rhs = refkat.linear([("a", "b")]).to_spec().to_oracle(O).freeze()
    assert O.csp_mutable(O.Mutable.compile_string(b"a"), rhs, 1).output_bytes() == b"b"


def test_bench_signatures_fixture(O):
    """SURVEY App. B signatures (an independent restatement) + the committed fixture."""
    fx = json.load(open(os.path.join(GOLD, "bench_signatures.json")))
    frozen = {}
    for row in fx["rows"]:
        if row["L"] > 100 and row["kind"] == 1:
            continue   # the large epsilon-dense rows run in the slow marker below
        key = (row["kind"], row["T"], row["B"])
        if key not in frozen:
            frozen[key] = O.Frozen.generate(*key)
        s = bytes(i % row["B"] for i in range(row["L"])) if row["kind"] == 0 else bytes(row["L"])
        p = O.csp_bytes(frozen[key], s)
        assert (p.tuples, p.relax_calls, len(p.ilabels), p.total, p.signature()) == \
            (row["N"], row["R"], row["P"], row["total"], row["sha16"]), row
        assert list(p.final_tuple) == row["final_tuple"]


def test_fuzz_fixture(O):
    fx = json.load(open(os.path.join(GOLD, "fuzz_paths.json")))
    for case in fx["cases"]:
        spec = Spec(case["num_states"], 0, case["finals"], [tuple(a) for a in case["arcs"]])
        f = spec.to_oracle(O).freeze()
        for s_hex, want in zip(case["strings"], case["paths"]):
            p = O.csp_bytes(f, bytes.fromhex(s_hex))
            got = None if p.status != O.STATUS_OK else [list(map(int, p.ilabels)), list(map(int, p.olabels)), [float(x) for x in p.weights], p.final_weight]
            assert got == want, (case["seed"], s_hex)


def test_generators_match_closed_forms(O):
    # ambiguous N(L) = 2L^2+3L+1, R(L) = 5 N(L-1) for 4L <= T
    f = O.Frozen.generate(O.KIND_AMBIGUOUS, 4096, 12)
    for L in (1, 5, 11, 19, 33):
        p = O.csp_bytes(f, bytes(L))
        assert p.tuples == 2 * L * L + 3 * L + 1
        assert p.relax_calls == 5 * (2 * (L - 1) ** 2 + 3 * (L - 1) + 1)
        assert len(p.ilabels) == L and p.total == 0.0


CYCLE_CASE = (4, [1.0, None, 0.0, None],
              [(3, 1, 1, 0.0, 0), (0, 0, 2, 0.0, 2), (3, 1, 1, 2.0, 3), (3, 2, 0, 1.0, 3), (2, 0, 0, 0.0, 0), (1, 0, 1, 1.0, 1),
               (0, 2, 0, 1.0, 1), (2, 2, 0, 1.0, 2), (2, 2, 1, 2.0, 3), (0, 2, 3, 2.0, 2), (2, 1, 1, 1.0, 3), (3, 0, 0, 2.0, 0),
               (1, 1, 0, 0.0, 1), (3, 2, 3, 2.0, 2), (0, 0, 3, 0.0, 2), (3, 2, 2, 0.0, 3), (2, 3, 1, 0.0, 0)], b"\x00\x00\x00")


def test_hazard_backtrack_cycle_is_flagged(O):
    # zero-weight epsilon cycles can make back[] cyclic (the reference then appends until
    # it runs out of memory, compose-shortest-path.zig:374-380); the oracle flags the case instead.
    n, finals, arcs, s = CYCLE_CASE
    f = Spec(n, 0, finals, arcs).to_oracle(O).freeze()
    assert O.csp_bytes(f, s).status == O.STATUS_BACKTRACK_CYCLE


def test_compose_bytes_lattice_dump(O):
    """The eager lattice dump used as the checker of fst_b200_compose_frozen_lattice_batch: compose.zig:29-198 on a
    compiled string — states in BFS discovery order, arcs of a state in frozen order (match arcs, then the
    transducer's input-epsilon arcs with ilabel 0), final weight fw1 (x) fw2."""
    import numpy as np
    from common import Spec
    # state 0: label 10 -> 1 (two parallel arcs, olabels 100 / 101), epsilon -> 2; state 1 final 0.5; state 2: label 10 -> 1
    rhs = Spec(3, 0, [None, 0.5, None], [(0, 10, 101, 2.0, 1), (0, 10, 100, 0.0, 1), (0, 0, 7, 1.0, 2), (2, 10, 102, 0.25, 1)])
    f = rhs.to_oracle(O).freeze()
    m = O.compose_bytes(f, bytes([9]))            # byte 9 = label 10
    start, ab, fin, il, ol, w, nx = m.dump()
    # tuples: 0 = (0, s0, f0); match arcs first -> (1, s1, f0) = 1; epsilon arc -> (0, s2, f1) = 2; from 2: match -> 1
    assert start == 0 and m.num_states() == 3
    assert list(ab) == [0, 3, 3, 4]
    assert list(il) == [10, 10, 0, 10] and list(ol) == [100, 101, 7, 102]
    assert list(w) == [0.0, 2.0, 1.0, 0.25] and list(nx) == [1, 1, 2, 1]
    assert np.isinf(fin[0]) and fin[1] == 0.5 and np.isinf(fin[2])
    # the empty string: one state, final only if the start state is final; epsilon arcs still expand
    m0 = O.compose_bytes(f, b"")
    s0, ab0, fin0, il0, ol0, w0, nx0 = m0.dump()
    assert m0.num_states() == 2 and list(il0) == [0] and list(ol0) == [7] and list(nx0) == [1] and np.isinf(fin0).all()
