"""GPU parity at the LITERAL sizes of BASELINE.json's configs and on the weight hazards of SURVEY App. A
("next round" item 1): the CUDA path through the C ABI against the CPU oracle, bit-exact."""
import hashlib
import json
import os
import random
import tempfile

import numpy as np
import pytest

from common import Spec, assert_batch_matches_oracle, frozen_pair, gen_image, random_rhs, random_string

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _sig(il, ol, w):
    return hashlib.sha256(";".join(f"{int(a)},{int(b)},{float(c)!r}" for a, b, c in zip(il, ol, w)).encode()).hexdigest()[:16]


@pytest.mark.parametrize("length", [96, 251])
def test_epsilon_dense_literal_lengths(L, O, gpu, length):
    """Config 2 (headline, len 96) and the survey's regression vector (len 251: P = 260, the arc order flips when
    (i + b) % 255 wraps, SURVEY App. B obs. v) at transducer-len 4096, branches 12: golden signature, N, R and the
    oracle's path bit for bit — through the batch entry (fast kernel and, forced, the general lean kernel)."""
    rows = [r for r in json.load(open(os.path.join(GOLD, "bench_signatures.json")))["rows"] if r["kind"] == 1 and r["L"] == length and r["T"] == 4096]
    assert rows, "golden row missing"
    row = rows[0]
    img = gen_image(O, 1, 4096, 12)
    forc, fprod = O.Frozen.from_bytes(img), L.Fst.from_image(img)
    s = bytes(length)
    p = O.csp_bytes(forc, s)
    assert p.signature()[:16] == row["sha16"] and len(p.ilabels) == row["P"]
    strings = [s, bytes(length - 1), s]
    try:
        for exhaustive in (1, 0):
            L.configure(exhaustive=exhaustive)
            res = assert_batch_matches_oracle(L, O, fprod, forc, strings)
            il, ol, w = res.path(0)
            assert _sig(il, ol, w) == row["sha16"] and len(il) == row["P"] and res.total(0) == row["total"]
            if exhaustive:
                assert res.n_tuples[0] == row["N"] == p.tuples
        os.environ["LIBFST_B200_NO_FAST"] = "1"
        L.configure()
        res = assert_batch_matches_oracle(L, O, L.Fst.from_image(img), forc, [s])
        il, ol, w = res.path(0)
        assert _sig(il, ol, w) == row["sha16"]
    finally:
        os.environ.pop("LIBFST_B200_NO_FAST", None)
        L.configure()


def test_ambiguous_and_plain_literal_lengths(L, O, gpu):
    """Config 1 / config 3 rows of the golden table at their literal sizes (ambiguous len 96 and 251, plain 251)."""
    names = {0: "plain", 2: "ambiguous"}
    from libfst_b200 import synth
    rows = [r for r in json.load(open(os.path.join(GOLD, "bench_signatures.json")))["rows"] if r["kind"] in names and r["T"] == 4096]
    assert len(rows) >= 4
    for kind in names:
        img = gen_image(O, kind, 4096, 12)
        forc, fprod = O.Frozen.from_bytes(img), L.Fst.from_image(img)
        mine = [r for r in rows if r["kind"] == kind]
        strings = [synth.input_string(names[kind], r["L"], r["B"]) for r in mine]
        try:
            for exhaustive in (1, 0):
                L.configure(exhaustive=exhaustive)
                res = assert_batch_matches_oracle(L, O, fprod, forc, strings)
                for i, r in enumerate(mine):
                    il, ol, w = res.path(i)
                    assert _sig(il, ol, w) == r["sha16"] and len(il) == r["P"] and res.total(i) == r["total"], r
                    if exhaustive:
                        assert res.n_tuples[i] == r["N"], r
        finally:
            L.configure()


def test_wetext_config4_literal_transducer(L, O, gpu):
    """Config 4 with the LITERAL transducer (K = 110 000 dictionary entries, ~1 M arcs; hash table, label index of the
    257-arc state, 8-record leader slab): 384 strings (lengths U[11,251], 70 % dictionary words) against the oracle
    through every table kind / lane count the engine can select for it."""
    from libfst_b200 import synth
    m, sources = synth.wetext_style(K=110000)
    fprod = m.freeze()
    with tempfile.NamedTemporaryFile(suffix=".fst", delete=False) as t:
        path = t.name
    try:
        assert fprod.save(path) == 0
        forc = O.Frozen.from_bytes(open(path, "rb").read())
    finally:
        os.unlink(path)
    assert fprod.num_states() > 700000 and sum(1 for _ in range(1)) == 1
    strings = synth.wetext_strings(sources, 384, seed=11)
    try:
        for engine, lanes in ((0, 0), (2, 8), (2, 16), (2, 32)):
            L.configure(engine=engine, lanes_per_string=lanes)
            res = assert_batch_matches_oracle(L, O, fprod, forc, strings if engine == 0 else strings[:96])
            assert (res.status == L.PATH).all()
        L.configure(exhaustive=1)
        res = assert_batch_matches_oracle(L, O, fprod, forc, strings[:64])
        for i in range(64):
            assert res.n_tuples[i] == O.csp_bytes(forc, strings[i]).tuples
    finally:
        L.configure()


def _rhs_with_infinities(rng, neg_inf):
    spec = random_rhs(rng, max_states=7)
    arcs = []
    for (src, il, ol, w, nxt) in spec.arcs:
        r = rng.random()
        if r < 0.18:
            w = float("inf")
        elif neg_inf and r < 0.24:
            w = float("-inf")
        arcs.append((src, il, ol, w, nxt))
    finals = list(spec.finals)
    if neg_inf:
        for s in range(len(finals)):
            if rng.random() < 0.1:
                finals[s] = float("-inf")
    return Spec(spec.num_states, 0, finals, arcs)


@pytest.mark.parametrize("neg_inf", [False, True])
def test_infinite_arc_weights(L, O, gpu, neg_inf):
    """SURVEY hazards H4 / H7 (weight.zig:19-32, compose-shortest-path.zig:109-114): arcs of weight +inf are still
    traversed (a brand-new target is always taken), -inf counts as Zero; the result can be an infinite-weight path."""
    rng = random.Random(20261018 + int(neg_inf))
    n_inf_paths = 0
    for case in range(150):
        spec = _rhs_with_infinities(rng, neg_inf)
        fprod, forc, _ = frozen_pair(L, O, spec)
        strings = [random_string(rng, max_len=7) for _ in range(24)]
        res = assert_batch_matches_oracle(L, O, fprod, forc, strings, check_out=True)
        for i in range(len(strings)):
            if res.status[i] == L.PATH and np.isinf(res.path(i)[2]).any():
                n_inf_paths += 1
        # the single-call drop-in too
        for s in strings[:3]:
            po = O.csp_bytes(forc, s)
            r = L.compose_frozen_shortest_path(L.MutableFst.compile_string(s), fprod, 1)
            if po.status == O.STATUS_BACKTRACK_CYCLE:
                assert r is None
            elif po.status == O.STATUS_EMPTY:
                assert r is not None and r.num_states() == 0
            else:
                il, ol, w, fw = r.chain()
                assert np.array_equal(il, po.ilabels) and np.array_equal(ol, po.olabels)
                assert np.array_equal(w.view(np.uint64), po.weights.view(np.uint64))
    assert n_inf_paths > 0   # the generator really produces infinite-weight results (H7)


def test_negative_final_weights(L, O, gpu):
    """Hazard H2 with non-negative arcs: negative FINAL weights only (the early stop must not apply)."""
    rng = random.Random(606)
    for case in range(200):
        spec = random_rhs(rng, max_states=8)
        spec.finals = [float(rng.randint(-3, 2)) if rng.random() < 0.5 else None for _ in range(spec.num_states)]
        fprod, forc, _ = frozen_pair(L, O, spec)
        strings = [random_string(rng, max_len=8) for _ in range(24)]
        assert_batch_matches_oracle(L, O, fprod, forc, strings)


def test_fuzz_100k_distinct_cases(L, O, gpu):
    """SURVEY 7.2 gate: >= 10^5 DISTINCT (transducer, string) cases of the tie-heavy generator (App. C), seeds fixed:
    2 600 random transducers x 40 distinct strings each."""
    rng = random.Random(7_2_2026)
    total = 0
    seen_specs = set()
    for case in range(2600):
        spec = random_rhs(rng, max_states=8, real=(case % 7 == 6))
        key = (spec.num_states, tuple(spec.finals), tuple(spec.arcs))
        if key in seen_specs:
            continue
        seen_specs.add(key)
        fprod, forc, _ = frozen_pair(L, O, spec)
        strings = set()
        while len(strings) < 40:
            strings.add(random_string(rng, max_len=9))
        strings = sorted(strings)
        assert_batch_matches_oracle(L, O, fprod, forc, strings, check_out=(case % 10 == 0))
        total += len(strings)
    assert total >= 100000


def test_empty_batch_first_call_and_odd_shapes(L, O, gpu):
    """ADVICE: n_strings == 0 must give FST_OK with an empty result (also as the very first call on a device), and
    batches of only empty strings work."""
    img = gen_image(O, 2, 64, 4)
    fprod, forc = L.Fst.from_image(img), O.Frozen.from_bytes(img)
    res = L.compose_frozen_shortest_path_batch(fprod, np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(res.status) == 0
    assert_batch_matches_oracle(L, O, fprod, forc, [b"", b"", b""])
    lat = L.compose_frozen_lattice_batch(fprod, np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(lat.status) == 0


def _reverse_ilabel_runs(image: bytes) -> bytes:
    """The same transducer with the arcs of every ilabel run of every state in reverse order (ilabels stay sorted)."""
    hdr = np.frombuffer(image, np.uint32, 6)
    ns, na = int(hdr[2]), int(hdr[3])
    st = np.frombuffer(image, np.dtype([("off", "<u4"), ("n", "<u4"), ("fw", "<f8")]), ns, 24)
    adt = np.dtype([("il", "<u4"), ("ol", "<u4"), ("w", "<f8"), ("nx", "<u4"), ("pad", "<u4")])
    arcs = np.frombuffer(image, adt, na, 24 + 16 * ns).copy()
    for s in range(ns):
        o, n = int(st["off"][s]), int(st["n"][s])
        i = 0
        while i < n:
            j = i
            while j < n and arcs["il"][o + j] == arcs["il"][o + i]:
                j += 1
            arcs[o + i:o + j] = arcs[o + i:o + j][::-1].copy()
            i = j
    return image[:24 + 16 * ns] + arcs.tobytes()


def test_unsorted_olabel_image(L, O, gpu):
    """ADVICE: fst_load / from_image accept any image whose ilabels are non-decreasing (fst.zig:227-273), also when
    the (olabel, weight, nextstate) order inside an ilabel run is not the freeze order.  The upload's parallel-arc
    fold and the back-track's first-tight-arc recovery assume the freeze order; such images must still match."""
    rng = random.Random(99)
    for case in range(120):
        spec = random_rhs(rng, max_states=6)
        img2 = _reverse_ilabel_runs(spec.to_oracle(O).freeze().to_bytes())
        fprod, forc = L.Fst.from_image(img2), O.Frozen.from_bytes(img2)
        strings = [random_string(rng, max_len=7) for _ in range(20)]
        assert_batch_matches_oracle(L, O, fprod, forc, strings)


def test_output_labels_above_256(L, O, gpu):
    """ADVICE: an output label above 256 has no byte form (string.zig:64-97: fst_print_output_string returns -1).  The
    batch entry reports FST_B200_NOT_BYTES with valid path arrays and an empty output string, and the pipeline entry
    does not feed such a string to its second stage."""
    rng = random.Random(300)
    seen = 0
    for case in range(40):
        spec = random_rhs(rng, max_states=6)
        spec.arcs = [(a, il, (ol + 255 if rng.random() < 0.3 and ol else ol), w, n) for (a, il, ol, w, n) in spec.arcs]
        fprod, forc, _ = frozen_pair(L, O, spec)
        strings = [random_string(rng, max_len=6) for _ in range(20)]
        res = assert_batch_matches_oracle(L, O, fprod, forc, strings)
        seen += int((res.status == L.NOT_BYTES).sum())
        data, offsets = L.pack_strings(strings)
        two = L.compose_frozen_shortest_path_pipeline(fprod, fprod, data, offsets)
        for i in range(len(strings)):
            if res.status[i] == L.NOT_BYTES:
                assert two.status[i] == L.NOT_BYTES and len(two.path(i)[0]) == 0
    assert seen > 0


def test_semantics_per_call_and_concurrent_callers(L, O, gpu):
    """VERDICT: (a) the eager pair is its own entry point — a thread using it and a thread using the lazy entry on the
    SAME frozen handle do not interfere; (b) the reference expects concurrent calls on one frozen handle
    (include/fst.h:11-26, src/c-api.zig:279-282) and defers fst_free while a call holds the handle
    (src/c-api.zig:220-248): the single call, the batch entry and fst_free race here."""
    import threading
    from common import assert_batch_matches_eager_oracle
    rng = random.Random(8)
    specs = [random_rhs(rng, max_states=7) for _ in range(6)]
    pairs = [frozen_pair(L, O, sp) for sp in specs]
    strings = [random_string(rng, max_len=8) for _ in range(64)]
    data, offsets = L.pack_strings(strings)
    errors = []

    def lazy_worker():
        try:
            for rep in range(12):
                fprod, forc, _ = pairs[rep % len(pairs)]
                assert_batch_matches_oracle(L, O, fprod, forc, strings)
        except Exception as e:   # noqa: BLE001
            errors.append(("lazy", repr(e)))

    def eager_worker():
        try:
            for rep in range(12):
                fprod, forc, _ = pairs[(rep + 1) % len(pairs)]
                res = L.compose_frozen_then_shortest_path_batch(fprod, data, offsets)
                assert_batch_matches_eager_oracle(L, O, fprod, forc, strings, res=res)
        except Exception as e:   # noqa: BLE001
            errors.append(("eager", repr(e)))

    def single_worker():
        try:
            for rep in range(40):
                fprod, forc, _ = pairs[rep % len(pairs)]
                s = strings[rep % len(strings)]
                po = O.csp_bytes(forc, s)
                r = L.compose_frozen_shortest_path(L.MutableFst.compile_string(s), fprod, 1)
                if po.status == O.STATUS_BACKTRACK_CYCLE:
                    assert r is None
                elif po.status == O.STATUS_EMPTY:
                    assert r is not None and r.num_states() == 0
                else:
                    il, ol, w, fw = r.chain()
                    assert np.array_equal(il, po.ilabels) and np.array_equal(ol, po.olabels)
        except Exception as e:   # noqa: BLE001
            errors.append(("single", repr(e)))

    ths = [threading.Thread(target=f) for f in (lazy_worker, eager_worker, single_worker, lazy_worker)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors

    # fst_free while searches are in flight: the handle dies only after the last call holding it returned
    img = gen_image(O, 1, 2048, 12)
    forc = O.Frozen.from_bytes(img)
    want = O.csp_bytes(forc, bytes(24))
    for rep in range(4):
        f = L.Fst.from_image(img)
        h = f.h
        d2, o2 = L.pack_strings([bytes(24)] * 8)
        got = []

        def searcher():
            try:
                res = L.compose_frozen_shortest_path_batch(f, d2, o2)
                got.append(res)
            except Exception as e:   # noqa: BLE001
                got.append(e)

        ts = [threading.Thread(target=searcher) for _ in range(3)]
        for t in ts:
            t.start()
        L.lib().fst_free(h)          # races with the searches: deferred while pinned
        f.h = L.FST_INVALID_HANDLE    # (the wrapper must not free it again)
        for t in ts:
            t.join()
        for g in got:
            if isinstance(g, Exception):
                assert "FstError 2" in str(g), g       # the call lost the race: the handle was already invalid
            else:
                il, ol, w = g.path(0)
                assert np.array_equal(il, want.ilabels) and np.array_equal(ol, want.olabels)
        assert L.lib().fst_num_states(h) == 0            # the handle is dead afterwards


@pytest.mark.parametrize("skew", ["-1", "0", "3", "7"])
def test_dense_table_layouts(L, O, gpu, skew):
    """The dense table's index map (a row per string position, or a row per diagonal state - skew * position) is a pure
    layout choice: every skew must give the oracle's results — fast kernel, general lean kernel, 16-byte records and
    the eager kernels.  LIBFST_B200_SKEW overrides the upload heuristic (read when a handle is first searched)."""
    from common import assert_batch_matches_eager_oracle
    rng = random.Random(31 + int(skew))
    img = gen_image(O, 1, 512, 12)
    forc = O.Frozen.from_bytes(img)
    strings = [bytes(n) for n in (33, 0, 1, 19, 40, 7)]
    specs = [random_rhs(rng, max_states=8, real=(k % 3 == 2)) for k in range(12)]
    rstrings = [random_string(rng, max_len=9) for _ in range(30)]
    os.environ["LIBFST_B200_SKEW"] = skew
    try:
        for no_fast in (False, True):
            if no_fast:
                os.environ["LIBFST_B200_NO_FAST"] = "1"
            for exhaustive in (0, 1):
                L.configure(exhaustive=exhaustive)
                assert_batch_matches_oracle(L, O, L.Fst.from_image(img), forc, strings)
            L.configure(engine=3, lanes_per_string=16)
            assert_batch_matches_oracle(L, O, L.Fst.from_image(img), forc, strings)
            L.configure()
            for spec in specs:
                fprod, fo, _ = frozen_pair(L, O, spec)
                assert_batch_matches_oracle(L, O, fprod, fo, rstrings)
        f = L.Fst.from_image(img)
        data, offsets = L.pack_strings(strings)
        assert_batch_matches_eager_oracle(L, O, f, forc, strings, res=L.compose_frozen_then_shortest_path_batch(f, data, offsets))
        # mixed lengths in one batch: every string uses the rows of the longest
        amb = gen_image(O, 2, 4096, 12)
        assert_batch_matches_oracle(L, O, L.Fst.from_image(amb), O.Frozen.from_bytes(amb), [bytes(n) for n in (96, 5, 64, 0, 33)])
    finally:
        os.environ.pop("LIBFST_B200_SKEW", None)
        os.environ.pop("LIBFST_B200_NO_FAST", None)
        L.configure()
