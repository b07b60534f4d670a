"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit-exact."""
import random

import numpy as np
import pytest

from common import (Spec, assert_batch_matches_oracle, frozen_pair, gen_image, random_lhs, random_rhs, random_string)

pytestmark = pytest.mark.gpu


def test_reference_unit_vector(L, O, gpu):
    # reference src/ops/compose-shortest-path.zig:447-471: "123" o ("123" -> "abc")
    rhs = O.Mutable.compile_string_transducer(b"123", b"abc").freeze()
    f = L.Fst.from_image(rhs.to_bytes())
    a = L.MutableFst.compile_string(b"123")
    r = L.compose_frozen_shortest_path(a, f, 1)
    assert r is not None and r.num_states() == 4 and r.start() == 0
    il, ol, w, fw = r.chain()
    assert list(il) == [50, 51, 52] and list(ol) == [98, 99, 100] and list(w) == [0.0, 0.0, 0.0] and fw == 0.0
    assert r.print_string(True) == b"abc" and r.print_string(False) == b"123"
    # n == 0 -> empty FST, n > 1 -> invalid handle (compose-shortest-path.zig:30-33)
    e = L.compose_frozen_shortest_path(a, f, 0)
    assert e is not None and e.num_states() == 0 and e.start() == L.FST_NO_STATE and e.print_string(True) is None
    assert L.compose_frozen_shortest_path(a, f, 2) is None


@pytest.mark.parametrize("kind,Lens", [(2, [11, 19, 33, 96]), (0, [11, 96, 251]), (1, [11])])
def test_bench_scenarios_small(L, O, gpu, kind, Lens):
    img = gen_image(O, kind, 4096, 12)
    forc = O.Frozen.from_bytes(img)
    fprod = L.Fst.from_image(img)
    if kind == 0:
        strings = [bytes(i % 12 for i in range(n)) for n in Lens]
    else:
        strings = [bytes(n) for n in Lens]
    res = assert_batch_matches_oracle(L, O, fprod, forc, strings)
    for i, s in enumerate(strings):
        p = O.csp_bytes(forc, s)
        assert res.n_tuples[i] <= p.tuples


def test_random_tie_heavy_batch(L, O, gpu):
    rng = random.Random(20261018)
    total = 0
    for case in range(60):
        spec = random_rhs(rng)
        fprod, forc, _ = frozen_pair(L, O, spec)
        strings = [random_string(rng) for _ in range(40)]
        assert_batch_matches_oracle(L, O, fprod, forc, strings)
        total += len(strings)
    assert total == 2400


def test_random_real_weights_batch(L, O, gpu):
    rng = random.Random(7)
    for case in range(30):
        spec = random_rhs(rng, max_states=12, real=True)
        fprod, forc, _ = frozen_pair(L, O, spec)
        strings = [random_string(rng, max_len=10) for _ in range(40)]
        assert_batch_matches_oracle(L, O, fprod, forc, strings)


def test_exhaustive_equals_early_exit(L, O, gpu):
    rng = random.Random(99)
    cases = []
    for case in range(20):
        spec = random_rhs(rng, max_states=10)
        fprod, forc, _ = frozen_pair(L, O, spec)
        strings = [random_string(rng, max_len=8) for _ in range(30)]
        cases.append((fprod, forc, strings))
    try:
        L.configure(exhaustive=1)
        for fprod, forc, strings in cases:
            res = assert_batch_matches_oracle(L, O, fprod, forc, strings)
            for i, s in enumerate(strings):
                assert res.n_tuples[i] == O.csp_bytes(forc, s).tuples   # literal: same number of tuples
    finally:
        L.configure(exhaustive=0)
    for fprod, forc, strings in cases:
        assert_batch_matches_oracle(L, O, fprod, forc, strings)


@pytest.mark.parametrize("lanes", [4, 8, 16, 32])
def test_group_sizes_agree(L, O, gpu, lanes):
    rng = random.Random(1234)
    try:
        L.configure(lanes_per_string=lanes)
        for case in range(15):
            spec = random_rhs(rng, max_states=10)
            fprod, forc, _ = frozen_pair(L, O, spec)
            strings = [random_string(rng, max_len=8) for _ in range(30)]
            assert_batch_matches_oracle(L, O, fprod, forc, strings)
        img = gen_image(O, 2, 512, 12)
        assert_batch_matches_oracle(L, O, L.Fst.from_image(img), O.Frozen.from_bytes(img), [bytes(33), bytes(5), b""])
    finally:
        L.configure(lanes_per_string=0)


def test_negative_weights_serial_mode(L, O, gpu):
    rng = random.Random(5)
    for case in range(40):
        spec = random_rhs(rng, neg=True, wmax=2)
        fprod, forc, _ = frozen_pair(L, O, spec)
        strings = [random_string(rng) for _ in range(20)]
        assert_batch_matches_oracle(L, O, fprod, forc, strings)


def _compare_single(L, O, lhs_spec, fprod, forc, n=1):
    po = O.csp_mutable(lhs_spec.to_oracle(O), forc, n)
    r = L.compose_frozen_shortest_path(lhs_spec.to_product(L), fprod, n)
    if po.status in (O.STATUS_BACKTRACK_CYCLE, O.STATUS_UNSUPPORTED_N):
        assert r is None
        return
    assert r is not None
    if po.status == O.STATUS_EMPTY:
        assert r.num_states() == 0
        return
    il, ol, w, fw = r.chain()
    assert np.array_equal(il, po.ilabels) and np.array_equal(ol, po.olabels)
    assert np.array_equal(w.view(np.uint64), po.weights.view(np.uint64))
    assert np.float64(fw).view(np.uint64) == np.float64(po.final_weight).view(np.uint64)
    assert r.num_states() == len(il) + 1


def test_general_lhs_single_call(L, O, gpu):
    rng = random.Random(31337)
    for case in range(150):
        spec = random_rhs(rng)
        fprod, forc, _ = frozen_pair(L, O, spec)
        for _ in range(3):
            _compare_single(L, O, random_lhs(rng), fprod, forc)


def test_general_lhs_negative(L, O, gpu):
    rng = random.Random(4242)
    for case in range(60):
        spec = random_rhs(rng, neg=True, wmax=2)
        fprod, forc, _ = frozen_pair(L, O, spec)
        lhs = random_lhs(rng, neg=True)
        _compare_single(L, O, lhs, fprod, forc)


def test_retry_passes_small_workspace(L, O, gpu):
    # force tiny arenas so the retry path (8x larger arenas, fewer groups) is exercised
    img = gen_image(O, 2, 4096, 12)
    forc, fprod = O.Frozen.from_bytes(img), L.Fst.from_image(img)
    strings = [bytes(n) for n in (5, 64, 7, 96, 3, 33)]
    try:
        L.configure(tuples_hint=64)
        res = assert_batch_matches_oracle(L, O, fprod, forc, strings)
        assert res.passes >= 2
    finally:
        L.configure()


def test_reference_kats_on_gpu(L, O, gpu):
    # rewrite.zig:244-498 KAT strings through the batched entry and the single-call drop-in
    import refkat
    for name, rule, cases in refkat.rewrite_kats():
        spec = rule.to_spec()
        fprod, forc, _ = frozen_pair(L, O, spec)
        strings = [c[0] for c in cases]
        res = assert_batch_matches_oracle(L, O, fprod, forc, strings)
        for i, (inp, want) in enumerate(cases):
            assert res.output(i) == want, (name, inp)
            r = L.compose_frozen_shortest_path(L.MutableFst.compile_string(inp), fprod, 1)
            assert r.print_string(True) == want and r.print_string(False) == inp


def test_golden_fixtures_on_gpu(L, O, gpu):
    import json
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden")
    fx = json.load(open(os.path.join(gold, "fuzz_paths.json")))
    for case in fx["cases"]:
        spec = Spec(case["num_states"], 0, case["finals"], [tuple(a) for a in case["arcs"]])
        fprod = spec.to_product(L).freeze()
        strings = [bytes.fromhex(h) for h in case["strings"]]
        data, offsets = L.pack_strings(strings)
        res = L.compose_frozen_shortest_path_batch(fprod, data, offsets)
        for i, want in enumerate(case["paths"]):
            if want is None:
                assert res.status[i] != L.PATH
                continue
            il, ol, w = res.path(i)
            assert [il.tolist(), ol.tolist(), w.tolist(), float(res.final_weights[i])] == want, (case["seed"], i)
    import hashlib
    fx = json.load(open(os.path.join(gold, "bench_signatures.json")))
    from libfst_b200 import synth
    names = {0: "plain", 1: "epsilon_dense", 2: "ambiguous"}
    for row in fx["rows"]:
        f = synth.TRANSDUCERS[names[row["kind"]]](row["T"], row["B"]).freeze()
        s = synth.input_string(names[row["kind"]], row["L"], row["B"])
        data, offsets = L.pack_strings([s])
        res = L.compose_frozen_shortest_path_batch(f, data, offsets)
        il, ol, w = res.path(0)
        sig = hashlib.sha256(";".join(f"{int(a)},{int(b)},{float(c)!r}" for a, b, c in zip(il, ol, w)).encode()).hexdigest()[:16]
        assert sig == row["sha16"] and len(il) == row["P"] and res.total(0) == row["total"], row


def test_mixed_lengths_large(L, O, gpu):
    # issue-#1 profile matrix lengths (run_issue1_profile_bench.py:24-25) in ONE batch: exercises retry passes + ordering
    img = gen_image(O, 2, 4096, 12)
    forc, fprod = O.Frozen.from_bytes(img), L.Fst.from_image(img)
    lens = [11, 19, 33, 64, 96, 128, 160, 192, 224, 251]
    rng = random.Random(1)
    strings = [bytes(rng.choice(lens)) for _ in range(300)]
    assert_batch_matches_oracle(L, O, fprod, forc, strings)
    try:
        L.configure(exhaustive=1)
        res = assert_batch_matches_oracle(L, O, fprod, forc, strings[:40])
        for i, s in enumerate(strings[:40]):
            n = len(s)
            assert res.n_tuples[i] == 2 * n * n + 3 * n + 1
    finally:
        L.configure()


def test_epsilon_dense_medium(L, O, gpu):
    img = gen_image(O, 1, 4096, 12)
    forc, fprod = O.Frozen.from_bytes(img), L.Fst.from_image(img)
    assert_batch_matches_oracle(L, O, fprod, forc, [bytes(33), bytes(19), bytes(11), bytes(64)])


def test_high_degree_states_use_search(L, O, gpu):
    # states with > 32 arcs exercise the G-ary equal_range narrowing and multi-chunk expansions
    rng = random.Random(77)
    n = 6
    arcs = []
    for s in range(n):
        for _ in range(rng.randint(60, 300)):
            il = 0 if rng.random() < 0.1 else rng.randint(1, 40)
            arcs.append((s, il, rng.randint(0, 5), float(rng.randint(0, 2)), rng.randrange(n)))
    spec = Spec(n, 0, [0.0 if rng.random() < 0.5 else None for _ in range(n)], arcs)
    fprod, forc, _ = frozen_pair(L, O, spec)
    strings = [bytes(rng.randint(0, 39) for _ in range(rng.randint(0, 5))) for _ in range(60)]
    for lanes in (0, 4, 32):
        L.configure(lanes_per_string=lanes)
        try:
            assert_batch_matches_oracle(L, O, fprod, forc, strings)
        finally:
            L.configure()


def test_cycle_hazard_is_reported(L, O, gpu):
    from test_oracle_golden import CYCLE_CASE
    n, finals, arcs, s = CYCLE_CASE
    fprod, forc, _ = frozen_pair(L, O, Spec(n, 0, finals, arcs))
    data, offsets = L.pack_strings([s, b""])
    res = L.compose_frozen_shortest_path_batch(fprod, data, offsets)
    assert res.status[0] == L.CYCLE
    assert L.compose_frozen_shortest_path(L.MutableFst.compile_string(s), fprod, 1) is None   # reference: OOM -> invalid handle


@pytest.mark.parametrize("engine,lanes", [(1, 0), (2, 32), (2, 16), (2, 8), (3, 32), (3, 16), (3, 8), (0, 0), (4, 0), (5, 0), (6, 0), (7, 0), (2, 4), (3, 4)])
def test_engines_agree(L, O, gpu, engine, lanes):
    """Every kernel choice for byte-string batches (general warp kernel, lean + hash table, lean + dense table;
    32 or 16 lanes per string) must reproduce the oracle bit for bit, in early-exit and exhaustive mode."""
    rng = random.Random(555 + engine)
    cases = []
    for case in range(25):
        spec = random_rhs(rng, real=(case % 3 == 2))
        fprod, forc, _ = frozen_pair(L, O, spec)
        cases.append((fprod, forc, [random_string(rng, max_len=9) for _ in range(40)]))
    # high-degree states: multi-chunk expansions and the binary-searched match range
    n = 5
    arcs = []
    for s in range(n):
        for _ in range(rng.randint(40, 200)):
            il = 0 if rng.random() < 0.1 else rng.randint(1, 30)
            arcs.append((s, il, rng.randint(0, 5), float(rng.randint(0, 2)), rng.randrange(n)))
    fprod, forc, _ = frozen_pair(L, O, Spec(n, 0, [0.0 if rng.random() < 0.5 else None for _ in range(n)], arcs))
    cases.append((fprod, forc, [bytes(rng.randint(0, 29) for _ in range(rng.randint(0, 5))) for _ in range(50)]))
    for kind, lens in ((2, [0, 1, 11, 33, 96]), (1, [0, 5, 11, 19]), (0, [7, 96])):
        img = gen_image(O, kind, 4096 if kind != 1 else 512, 12)
        strings = [bytes(i % 12 for i in range(k)) if kind == 0 else bytes(k) for k in lens]
        cases.append((L.Fst.from_image(img), O.Frozen.from_bytes(img), strings))
    try:
        for exhaustive in (0, 1):
            L.configure(engine=engine, lanes_per_string=lanes, exhaustive=exhaustive)
            for fprod, forc, strings in cases:
                res = assert_batch_matches_oracle(L, O, fprod, forc, strings)
                if exhaustive:
                    for i, s in enumerate(strings):
                        p = O.csp_bytes(forc, s)
                        if p.status == O.STATUS_OK:
                            assert res.n_tuples[i] == p.tuples
    finally:
        L.configure()


def test_lean_window_eviction_and_levels(L, O, gpu):
    """Searches with more than one ready-bitmap line (> 1024 ids), many distance levels and queue jumps
    (an old tuple lowered to the current level below the window) on both lean tables."""
    rng = random.Random(2026)
    n = 60
    arcs = []
    for s in range(n):
        for _ in range(rng.randint(2, 6)):
            il = 0 if rng.random() < 0.25 else rng.randint(1, 2)
            arcs.append((s, il, rng.randint(0, 3), float(rng.choice([0, 0, 0, 1, 2, 5])), rng.randrange(n)))
    fprod, forc, _ = frozen_pair(L, O, Spec(n, 0, [0.0 if rng.random() < 0.3 else None for _ in range(n)], arcs))
    strings = [bytes(rng.randint(0, 1) for _ in range(rng.randint(20, 60))) for _ in range(24)]
    try:
        for engine, lanes in ((2, 32), (3, 32), (3, 16), (2, 16), (3, 8), (2, 8), (0, 0), (5, 0), (6, 0), (3, 4), (2, 4)):
            for exhaustive in (1, 0):
                L.configure(engine=engine, lanes_per_string=lanes, exhaustive=exhaustive)
                res = assert_batch_matches_oracle(L, O, fprod, forc, strings)
                if exhaustive:
                    assert max(res.n_tuples) > 1100
    finally:
        L.configure()


def test_wetext_style_config4(L, O, gpu):
    """BASELINE config 4 at reduced scale: trie-shaped tagger with epsilon chains, fractional weights (many distance
    levels), one 257-arc state; every kernel choice against the oracle."""
    import os
    import tempfile
    from libfst_b200 import synth
    m, sources = synth.wetext_style(K=3000)
    fprod = m.freeze()
    with tempfile.NamedTemporaryFile(suffix=".fst", delete=False) as t:
        path = t.name
    try:
        assert fprod.save(path) == 0
        forc = O.Frozen.from_bytes(open(path, "rb").read())
    finally:
        os.unlink(path)
    strings = synth.wetext_strings(sources, 160, seed=3, lo=0, hi=80)
    try:
        for engine, lanes in ((0, 0), (2, 4), (2, 8), (2, 16), (2, 32), (1, 0)):
            L.configure(engine=engine, lanes_per_string=lanes)
            res = assert_batch_matches_oracle(L, O, fprod, forc, strings)
            assert (res.status == L.PATH).all()
        L.configure(exhaustive=1)
        assert_batch_matches_oracle(L, O, fprod, forc, strings[:40])
        # the optional 4-record leader slab (4 lanes per string picked automatically for a sparse transducer): the
        # device image is built when the transducer is first searched, so a fresh handle sees the switch
        os.environ["LIBFST_B200_LANES4"] = "1"
        f4 = m.freeze()
        L.configure()
        res = assert_batch_matches_oracle(L, O, f4, forc, strings)
        assert (res.status == L.PATH).all()
    finally:
        os.environ.pop("LIBFST_B200_LANES4", None)
        L.configure()


@pytest.mark.parametrize("engine,lanes", [(0, 0), (2, 32), (3, 16), (2, 8), (3, 8), (3, 4)])
def test_eager_semantics_config5(L, O, gpu, engine, lanes):
    """BASELINE config 5 / SURVEY rows a14+a15: the path of compose() followed by shortestPath() (different
    tie-breaking than the lazy search: lattice states numbered in FIFO order, no label tie-break)."""
    from common import assert_batch_matches_eager_oracle
    rng = random.Random(909 + engine)
    cases = []
    for case in range(40):
        spec = random_rhs(rng, real=(case % 4 == 3))
        fprod, forc, _ = frozen_pair(L, O, spec)
        cases.append((fprod, forc, [random_string(rng, max_len=8) for _ in range(30)]))
    for kind, lens in ((2, [0, 1, 11, 33, 96]), (1, [0, 5, 11]), (0, [7, 96])):
        img = gen_image(O, kind, 4096 if kind != 1 else 512, 12)
        strings = [bytes(i % 12 for i in range(k)) if kind == 0 else bytes(k) for k in lens]
        cases.append((L.Fst.from_image(img), O.Frozen.from_bytes(img), strings))
    differ = 0
    try:
        for exhaustive in (0, 1):
            L.configure(engine=engine, lanes_per_string=lanes, exhaustive=exhaustive, semantics=L.EAGER)
            for fprod, forc, strings in cases:
                assert_batch_matches_eager_oracle(L, O, fprod, forc, strings)
        # the two semantics really differ on some tie-heavy inputs (~2 %)
        L.configure(engine=engine, lanes_per_string=lanes, semantics=L.LAZY)
        for fprod, forc, strings in cases[:40]:
            data, offsets = L.pack_strings(strings)
            lazy = L.compose_frozen_shortest_path_batch(fprod, data, offsets)
            for i, s in enumerate(strings):
                pe, _, _ = O.eager_mutable(O.Mutable.compile_string(s), forc, 1)
                if pe.status == O.STATUS_OK and lazy.status[i] == L.PATH:
                    il, ol, w = lazy.path(i)
                    differ += not (np.array_equal(il, pe.ilabels) and np.array_equal(ol, pe.olabels))
    finally:
        L.configure()
    assert differ > 0


@pytest.mark.parametrize("engine", [0, 2])
def test_eager_config5_full_length(L, O, gpu, engine):
    """Config 5 at its literal size (label 1 x 251 against the ambiguous chain: 126 756 lattice states).  With the
    hash table (engine 2) the adaptive passes go through a capacity where the search phase nearly fills the table
    and the BFS phase meets new tuples on top of it — the table must report overflow, not fill up."""
    from common import assert_batch_matches_eager_oracle
    img = gen_image(O, 2, 4096, 12)
    fprod, forc = L.Fst.from_image(img), O.Frozen.from_bytes(img)
    strings = [bytes(251), bytes(250), bytes(160), bytes(251)]   # byte 0 = label 1
    try:
        L.configure(engine=engine, semantics=L.EAGER)
        res = assert_batch_matches_eager_oracle(L, O, fprod, forc, strings)
        assert res.n_tuples[0] == 126756
    finally:
        L.configure()


@pytest.mark.parametrize("engine", [5, 6])
def test_wave_kernel_tie_heavy(L, O, gpu, engine):
    """Wave kernel (one warp per string, a ready word of up to 32 tuples per step, csp_wave.cuh): wide ready sets
    with massive exact ties, several distance levels, tuples created above the level and lowered inside the same
    chunk, existing tuples lowered to the level (chunk abandoned -> single pop), states wider than the wave slab."""
    rng = random.Random(4242 + engine)
    cases = []
    for case in range(30):
        n = rng.randint(8, 120)
        nlab = rng.randint(1, 3)
        wchoice = rng.choice([[0], [0, 0, 0, 1], [0, 1, 2, 3], [0, 0.5, 0.25, 1.5], [0, 0, 0, 0, 5]])
        arcs = []
        for s in range(n):
            deg = rng.randint(1, 7) if case % 5 else rng.randint(1, 14)   # every fifth case has states wider than 8 records
            for _ in range(deg):
                il = 0 if rng.random() < 0.25 else rng.randint(1, nlab)
                nxt = min(n - 1, s + rng.randint(0, 4)) if rng.random() < 0.8 else rng.randrange(n)
                arcs.append((s, il, rng.randint(0, 3), float(rng.choice(wchoice)), nxt))
        finals = [float(rng.choice(wchoice)) if rng.random() < 0.3 else None for _ in range(n)]
        finals[n - 1] = 0.0
        fprod, forc, _ = frozen_pair(L, O, Spec(n, 0, finals, arcs))
        strings = [bytes(rng.randint(0, nlab - 1) for _ in range(rng.randint(0, 48))) for _ in range(24)]
        cases.append((fprod, forc, strings))
    try:
        for exhaustive in (1, 0):
            L.configure(engine=engine, exhaustive=exhaustive)
            for fprod, forc, strings in cases:
                res = assert_batch_matches_oracle(L, O, fprod, forc, strings)
                if exhaustive:
                    for i, s in enumerate(strings):
                        p = O.csp_bytes(forc, s)
                        if p.status == O.STATUS_OK:
                            assert res.n_tuples[i] == p.tuples
    finally:
        L.configure()


def test_wave_kernel_bench_shapes(L, O, gpu):
    """The bench transducers through the wave kernel at sizes the oracle finishes quickly (dense and hash table)."""
    try:
        for engine in (6, 5):
            L.configure(engine=engine)
            for kind, lens in ((1, [0, 1, 5, 11, 19, 33]), (2, [0, 1, 11, 33, 96, 160]), (0, [7, 96])):
                img = gen_image(O, kind, 4096 if kind != 1 else 1024, 12)
                strings = [bytes(i % 12 for i in range(k)) if kind == 0 else bytes(k) for k in lens]
                assert_batch_matches_oracle(L, O, L.Fst.from_image(img), O.Frozen.from_bytes(img), strings)
    finally:
        L.configure()


def test_compact_records_and_wide_retry(L, O, gpu):
    """Integer-weight transducers use 8-byte table records (dist:20 | id:22 | prev:22) in the dense lean kernel; a
    distance beyond 20 bits sends the string back with 16-byte records (status stays PATH, output identical).
    LIBFST_B200_NO_CREC=1 forces 16-byte records: both forms must agree with the oracle."""
    import os
    rng = random.Random(77)
    n = 330
    arcs = []
    for s in range(n - 1):
        arcs.append((s, 1, 1, 4000.0, s + 1))
        arcs.append((s, 1, 2, 4095.0, s + 1))
        if s % 3 == 0:
            arcs.append((s, 0, 3, 7.0, s + 1))
        arcs.append((s, 2, 2, float(rng.randint(0, 3)), min(n - 1, s + rng.randint(1, 2))))
    fprod, forc, _ = frozen_pair(L, O, Spec(n, 0, [0.0] * n, arcs))
    strings = [bytes(300), bytes(10), bytes([1] * 200), bytes(rng.randint(0, 1) for _ in range(280)), b""]
    try:
        for env in ("0", "1"):
            os.environ["LIBFST_B200_NO_CREC"] = env
            if env == "0":
                del os.environ["LIBFST_B200_NO_CREC"]
            for engine, lanes in ((3, 8), (3, 16), (0, 0)):
                L.configure(engine=engine, lanes_per_string=lanes)
                res = assert_batch_matches_oracle(L, O, fprod, forc, strings)
                assert res.final_weights[0] == 0.0 and res.status[0] == L.PATH
                il, ol, w = res.path(0)
                assert float(np.sum(w)) > 1048575.0      # the first string's distance really outgrows 20 bits
    finally:
        os.environ.pop("LIBFST_B200_NO_CREC", None)
        L.configure()


def test_two_stage_pipeline(L, O, gpu):
    """tagger then verbalizer on the device == two oracle searches chained through the output tape
    (compile_string -> composeShortestPath -> printStringFromTape -> compile_string -> composeShortestPath)."""
    rng = random.Random(31337)
    for case in range(12):
        a = random_rhs(rng, max_states=7, nlab=3, real=(case % 3 == 1))
        b = random_rhs(rng, max_states=7, nlab=3, real=(case % 3 == 2))
        fa, oa, _ = frozen_pair(L, O, a)
        fb, ob, _ = frozen_pair(L, O, b)
        strings = [random_string(rng, nlab=3, max_len=10) for _ in range(60)]
        data, offsets = L.pack_strings(strings)
        res = L.compose_frozen_shortest_path_pipeline(fa, fb, data, offsets)
        n_ok = 0
        for i, s in enumerate(strings):
            p1 = O.csp_bytes(oa, s)
            if p1.status == O.STATUS_BACKTRACK_CYCLE:
                assert res.status[i] == L.CYCLE, (case, i, s)
                continue
            if p1.status == O.STATUS_EMPTY:
                assert res.status[i] == L.NO_PATH and len(res.path(i)[0]) == 0, (case, i, s, res.status[i])
                continue
            mid = p1.output_bytes()
            # output labels of random transducers are 1..nlab -> bytes 0..nlab-1: valid second-stage inputs
            p2 = O.csp_bytes(ob, mid)
            if p2.status == O.STATUS_BACKTRACK_CYCLE:
                assert res.status[i] == L.CYCLE, (case, i, s)
                continue
            if p2.status == O.STATUS_EMPTY:
                assert res.status[i] == L.NO_PATH, (case, i, s, mid, res.status[i])
                continue
            assert res.status[i] == L.PATH, (case, i, s, mid, res.status[i])
            il, ol, w = res.path(i)
            assert np.array_equal(il, p2.ilabels) and np.array_equal(ol, p2.olabels), (case, i, s, mid)
            assert np.array_equal(w.view(np.uint64), p2.weights.view(np.uint64))
            assert np.float64(res.final_weights[i]).view(np.uint64) == np.float64(p2.final_weight).view(np.uint64)
            assert res.output(i) == p2.output_bytes()
            n_ok += 1
        assert n_ok > 0 or case > 0
    # the bench transducers chained: plain then plain (deterministic relabelling twice)
    img = gen_image(O, 0, 512, 12)
    f, o = L.Fst.from_image(img), O.Frozen.from_bytes(img)
    strings = [bytes(i % 12 for i in range(k)) for k in (0, 1, 7, 40)]
    data, offsets = L.pack_strings(strings)
    res = L.compose_frozen_shortest_path_pipeline(f, f, data, offsets)
    for i, s in enumerate(strings):
        p1 = O.csp_bytes(o, s)
        p2 = O.csp_bytes(o, p1.output_bytes()) if p1.status == O.STATUS_OK else None
        if p2 is not None and p2.status == O.STATUS_OK:
            assert res.status[i] == L.PATH and res.output(i) == p2.output_bytes()
        else:
            assert res.status[i] == L.NO_PATH


def test_eager_lattice_csr(L, O, gpu):
    """the lattice of compose(compile_string(s), b) as CSR from the device == the oracle's compose()
    (compose.zig:29-198): same state numbering (BFS discovery order), same arcs in the same order, same final weights."""
    rng = random.Random(2718)
    cases = []
    for case in range(20):
        spec = random_rhs(rng, max_states=9, real=(case % 4 == 3))
        fprod, forc, _ = frozen_pair(L, O, spec)
        cases.append((fprod, forc, [random_string(rng, max_len=9) for _ in range(25)]))
    for kind, lens in ((2, [0, 1, 11, 33]), (1, [0, 5, 11]), (0, [7, 40])):
        img = gen_image(O, kind, 4096 if kind != 1 else 512, 12)
        strings = [bytes(i % 12 for i in range(k)) if kind == 0 else bytes(k) for k in lens]
        cases.append((L.Fst.from_image(img), O.Frozen.from_bytes(img), strings))
    checked = 0
    for fprod, forc, strings in cases:
        data, offsets = L.pack_strings(strings)
        res = L.compose_frozen_lattice_batch(fprod, data, offsets)
        for i, s in enumerate(strings):
            m = O.compose_bytes(forc, s)
            start, ab, fin, il, ol, w, nx = m.dump()
            if m.num_states() == 0:
                assert res.status[i] == L.NO_PATH, (i, s, res.status[i])
                continue
            assert res.status[i] == L.PATH and start == 0, (i, s, res.status[i])
            gab, gfin, gil, gol, gw, gnx = res.lattice(i)
            assert np.array_equal(gab, ab), (i, s, gab[:8], ab[:8])
            assert np.array_equal(gfin.view(np.uint64), fin.view(np.uint64)), (i, s)
            assert np.array_equal(gil, il) and np.array_equal(gol, ol) and np.array_equal(gnx, nx), (i, s)
            assert np.array_equal(gw.view(np.uint64), w.view(np.uint64)), (i, s)
            checked += 1
    assert checked > 300


def test_large_batch_longest_first_order(L, O, gpu):
    """Batches of 4096+ strings leave the work queue longest first (device radix sort by length): results must still
    come back in input order, bit-exact, for strings of very different lengths (including empty ones)."""
    rng = random.Random(99)
    spec = random_rhs(rng, max_states=7, nlab=3)
    fprod, forc, _ = frozen_pair(L, O, spec)
    strings = [random_string(rng, nlab=3, max_len=rng.choice([0, 1, 3, 12, 30])) for _ in range(5000)]
    data, offsets = L.pack_strings(strings)
    res = L.compose_frozen_shortest_path_batch(fprod, data, offsets)
    cache = {}
    for i, s in enumerate(strings):
        p = cache.get(s)
        if p is None:
            p = cache[s] = O.csp_bytes(forc, s)
        if p.status == O.STATUS_BACKTRACK_CYCLE:
            assert res.status[i] == L.CYCLE
        elif p.status == O.STATUS_EMPTY:
            assert res.status[i] == L.NO_PATH
        else:
            assert res.status[i] == L.PATH, (i, s)
            il, ol, w = res.path(i)
            assert np.array_equal(il, p.ilabels) and np.array_equal(ol, p.olabels), (i, s)
            assert np.array_equal(w.view(np.uint64), p.weights.view(np.uint64)), (i, s)
            assert res.output(i) == p.output_bytes(), (i, s)
    # the same batch through the two-stage entry (stage 2 sorts stage 1's output strings)
    res2 = L.compose_frozen_shortest_path_pipeline(fprod, fprod, data, offsets)
    for i in range(0, 5000, 37):
        p1 = cache[strings[i]]
        if p1.status != O.STATUS_OK:
            assert res2.status[i] != L.PATH
            continue
        p2 = O.csp_bytes(forc, p1.output_bytes())
        if p2.status == O.STATUS_OK:
            assert res2.status[i] == L.PATH and res2.output(i) == p2.output_bytes(), (i, strings[i])


def test_length_segments(L, O, gpu):
    """With LIBFST_B200_SEGMENTS=1, batches of 32 768+ strings whose lengths differ by more than 4:3 are searched in length
    segments, each with its own arena geometry (LIBFST_B200_DEBUG shows them); every string must still match the
    oracle, in input order, with both table kinds (and without the switch)."""
    import os
    rng = random.Random(1234)
    spec = random_rhs(rng, max_states=6, nlab=2)
    fprod, forc, _ = frozen_pair(L, O, spec)
    lens = [0, 3, 10, 20, 40]
    strings = [bytes(rng.randint(0, 1) for _ in range(lens[i % 5])) for i in range(40000)]
    rng.shuffle(strings)
    data, offsets = L.pack_strings(strings)
    cache = {}
    try:
        for engine, seg in ((0, "1"), (2, "1"), (0, None)):
            if seg:
                os.environ["LIBFST_B200_SEGMENTS"] = seg
            else:
                os.environ.pop("LIBFST_B200_SEGMENTS", None)
            L.configure(engine=engine)
            res = L.compose_frozen_shortest_path_batch(fprod, data, offsets)
            for i, s in enumerate(strings):
                p = cache.get(s)
                if p is None:
                    p = cache[s] = O.csp_bytes(forc, s)
                if p.status == O.STATUS_BACKTRACK_CYCLE:
                    assert res.status[i] == L.CYCLE
                elif p.status == O.STATUS_EMPTY:
                    assert res.status[i] == L.NO_PATH, (i, s)
                else:
                    assert res.status[i] == L.PATH, (i, s, res.status[i])
                    il, ol, w = res.path(i)
                    assert np.array_equal(il, p.ilabels) and np.array_equal(ol, p.olabels), (i, s)
                    assert np.array_equal(w.view(np.uint64), p.weights.view(np.uint64)), (i, s)
    finally:
        os.environ.pop("LIBFST_B200_SEGMENTS", None)
        L.configure()
