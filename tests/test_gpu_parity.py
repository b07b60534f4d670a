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
