"""One batch sharded over the GPUs of the box through fst_compose_frozen_shortest_path_batch_multi (SURVEY 8e,
BASELINE north_star (3)): per-string results must not depend on the number of GPUs — 1-, 2-, 4- and 8-GPU outputs
byte-identical to the single-GPU batch entry (and to the oracle on a sample), in input order.

On a 1-GPU box the device lists collapse to [0]; the chunking / ordering / merge logic is still exercised
(several chunks on one device).  tests/test_sharding_gloo.py covers the torchrun harness of bench.py on CPU."""
import random

import numpy as np
import pytest

from common import assert_batch_matches_oracle, frozen_pair, gen_image, random_rhs, random_string

pytestmark = pytest.mark.gpu


def _same(a, b):
    assert np.array_equal(a.status, b.status)
    assert np.array_equal(a.path_offsets, b.path_offsets) and np.array_equal(a.out_offsets, b.out_offsets)
    assert np.array_equal(a.ilabels, b.ilabels) and np.array_equal(a.olabels, b.olabels)
    assert np.array_equal(a.weights.view(np.uint64), b.weights.view(np.uint64))
    assert np.array_equal(a.final_weights.view(np.uint64), b.final_weights.view(np.uint64))
    assert np.array_equal(a.out_bytes, b.out_bytes)


def _device_lists(L):
    n = L.device_count()
    return [list(range(k)) for k in (1, 2, 4, 8) if k <= n] or [[0]]


def test_multi_gpu_outputs_identical_mixed_lengths(L, O, gpu):
    """Mixed-length batch (issue-#1 lengths) on the ambiguous bench transducer + a tie-heavy random transducer."""
    rng = random.Random(5)
    img = gen_image(O, 2, 4096, 12)
    forc, fprod = O.Frozen.from_bytes(img), L.Fst.from_image(img)
    lens = [11, 19, 33, 64, 96, 128, 160, 192, 224, 251]
    strings = [bytes(rng.choice(lens)) for _ in range(6000)]
    data, offsets = L.pack_strings(strings)
    one = L.compose_frozen_shortest_path_batch(fprod, data, offsets)
    assert (one.status == L.PATH).all()
    for devs in _device_lists(L):
        for cpd in (1, 3):
            m = L.compose_frozen_shortest_path_batch_multi(fprod, data, offsets, devices=devs, chunks_per_device=cpd)
            assert m.n_devices == len(devs) and int(m.chunk_first[-1]) == len(strings)
            assert len(m.chunks) == min(len(devs) * cpd, len(strings))
            assert set(m.chunk_device.tolist()) <= set(devs)
            _same(m.flat(), one)
    assert_batch_matches_oracle(L, O, fprod, forc, strings[:40], res=None)

    spec = random_rhs(rng, max_states=7, nlab=3)
    fp2, fo2, _ = frozen_pair(L, O, spec)
    strings = [random_string(rng, nlab=3, max_len=rng.choice([0, 1, 3, 12, 30])) for _ in range(9000)]
    data, offsets = L.pack_strings(strings)
    one = L.compose_frozen_shortest_path_batch(fp2, data, offsets)
    for devs in _device_lists(L):
        m = L.compose_frozen_shortest_path_batch_multi(fp2, data, offsets, devices=devs)
        _same(m.flat(), one)
    # every visible device by default; and the oracle on a sample of the merged result
    m = L.compose_frozen_shortest_path_batch_multi(fp2, data, offsets)
    flat = m.flat()
    _same(flat, one)
    idx = list(range(0, 9000, 131))
    sub = [strings[i] for i in idx]
    d2, o2 = L.pack_strings(sub)
    assert_batch_matches_oracle(L, O, fp2, fo2, sub, res=L.compose_frozen_shortest_path_batch(fp2, d2, o2))
    for k, i in enumerate(idx):
        p = O.csp_bytes(fo2, strings[i])
        if p.status == O.STATUS_OK:
            il, ol, w = flat.path(i)
            assert np.array_equal(il, p.ilabels) and np.array_equal(ol, p.olabels)


def test_multi_gpu_edge_cases(L, O, gpu):
    img = gen_image(O, 1, 512, 12)
    fprod = L.Fst.from_image(img)
    # empty batch, fewer strings than devices x chunks, a repeated / unknown device
    m = L.compose_frozen_shortest_path_batch_multi(fprod, np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(m.chunks) == 0 and len(m.flat().status) == 0
    data, offsets = L.pack_strings([bytes(5), bytes(9), b""])
    one = L.compose_frozen_shortest_path_batch(fprod, data, offsets)
    m = L.compose_frozen_shortest_path_batch_multi(fprod, data, offsets, chunks_per_device=8)
    _same(m.flat(), one)
    with pytest.raises(RuntimeError):
        L.compose_frozen_shortest_path_batch_multi(fprod, data, offsets, devices=[0, 0])
    with pytest.raises(RuntimeError):
        L.compose_frozen_shortest_path_batch_multi(fprod, data, offsets, devices=[L.device_count()])


def test_multi_gpu_epsilon_dense_sample(L, O, gpu):
    """The headline workload split over the box's GPUs: identical strings, identical outputs, all devices used."""
    img = gen_image(O, 1, 4096, 12)
    forc, fprod = O.Frozen.from_bytes(img), L.Fst.from_image(img)
    n = 64 * max(1, L.device_count())
    strings = [bytes(33)] * n
    data, offsets = L.pack_strings(strings)
    m = L.compose_frozen_shortest_path_batch_multi(fprod, data, offsets, chunks_per_device=1)
    flat = m.flat()
    p = O.csp_bytes(forc, bytes(33))
    for i in range(0, n, 17):
        il, ol, w = flat.path(i)
        assert np.array_equal(il, p.ilabels) and np.array_equal(ol, p.olabels)
        assert np.array_equal(w.view(np.uint64), p.weights.view(np.uint64))
    assert (flat.status == L.PATH).all() and len(set(m.chunk_device.tolist())) == min(L.device_count(), len(m.chunks))


def test_result_without_path_arrays(L, O, gpu):
    """FST_B200_RESULT_NO_PATHS: output strings, statuses, final weights and path LENGTHS as usual, no per-arc arrays."""
    rng = random.Random(12)
    spec = random_rhs(rng, max_states=7, nlab=3)
    fprod, forc, _ = frozen_pair(L, O, spec)
    strings = [random_string(rng, nlab=3, max_len=12) for _ in range(500)]
    data, offsets = L.pack_strings(strings)
    full = L.compose_frozen_shortest_path_batch(fprod, data, offsets)
    lean = L.compose_frozen_shortest_path_batch(fprod, data, offsets, flags=L.RESULT_NO_PATHS)
    assert len(lean.ilabels) == 0 and len(lean.weights) == 0
    assert np.array_equal(lean.status, full.status) and np.array_equal(lean.path_offsets, full.path_offsets)
    assert np.array_equal(lean.out_offsets, full.out_offsets) and np.array_equal(lean.out_bytes, full.out_bytes)
    assert np.array_equal(lean.final_weights.view(np.uint64), full.final_weights.view(np.uint64))
    m = L.compose_frozen_shortest_path_batch_multi(fprod, data, offsets, chunks_per_device=3, flags=L.RESULT_NO_PATHS)
    for c, lo in zip(m.chunks, m.chunk_first[:-1]):
        lo = int(lo)
        for i in range(len(c.status)):
            assert c.status[i] == full.status[lo + i] and c.output(i) == full.output(lo + i)
    with pytest.raises(RuntimeError):
        L.compose_frozen_shortest_path_batch(fprod, data, offsets, flags=6)
