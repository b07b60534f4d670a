"""CPU: host-side logic of the product library and the C-ABI surface (no compute calls)."""
import ctypes as C
import os
import re
import struct

import pytest

from common import Spec, frozen_pair, random_rhs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "libfst_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(fst_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 35
    lib = C.CDLL(L._build.SO)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    # and the python binding covers every one of them
    assert names == set(L.EXPORTS), names ^ set(L.EXPORTS)


def test_stale_handle_cannot_touch_reused_slot(L):
    # reference src/c-api.zig:1426-1437
    lib = L.lib()
    first = lib.fst_mutable_new()
    assert first != L.FST_INVALID_HANDLE
    lib.fst_mutable_free(first)
    second = lib.fst_mutable_new()
    assert second != L.FST_INVALID_HANDLE and second != first
    assert lib.fst_mutable_add_state(first) == L.FST_NO_STATE
    assert lib.fst_mutable_add_state(second) == 0
    lib.fst_mutable_free(second)
    lib.fst_mutable_free(second)   # double free is a no-op
    # a mutable handle is never valid as a frozen one (separate tables, c-api.zig:276-277)
    m = lib.fst_mutable_new()
    assert lib.fst_num_states(m) == 0 and lib.fst_start(m) == L.FST_NO_STATE
    lib.fst_mutable_free(m)


def test_builder_error_codes(L):
    lib = L.lib()
    m = lib.fst_mutable_new()
    assert lib.fst_mutable_set_start(m, 0) == L.FST_INVALID_STATE          # c-api.zig:477
    assert lib.fst_mutable_add_state(m) == 0 and lib.fst_mutable_add_state(m) == 1
    assert lib.fst_mutable_set_start(m, 0) == L.FST_OK
    assert lib.fst_mutable_set_final(m, 5, 0.0) == L.FST_INVALID_STATE     # c-api.zig:486
    assert lib.fst_mutable_add_arc(m, 0, 1, 1, 0.5, 2) == L.FST_INVALID_STATE   # nextstate out of range, :496
    assert lib.fst_mutable_add_arc(m, 0, 1, 1, 0.5, 1) == L.FST_OK
    assert lib.fst_mutable_set_start(L.FST_INVALID_HANDLE, 0) == L.FST_INVALID_ARG
    assert lib.fst_mutable_num_arcs(m, 0) == 1 and lib.fst_mutable_num_arcs(m, 9) == 0
    assert lib.fst_mutable_final_weight(m, 0) == float("inf") and lib.fst_mutable_final_weight(m, 9) == float("inf")
    c = lib.fst_mutable_clone(m)
    assert c != L.FST_INVALID_HANDLE and lib.fst_mutable_num_states(c) == 2
    lib.fst_mutable_free(c)
    lib.fst_mutable_free(m)
    assert lib.fst_mutable_clone(m) == L.FST_INVALID_HANDLE


def test_freeze_matches_oracle_bytes(L, O):
    """fst_freeze (sort by (ilabel, olabel, weight, nextstate), pack) is byte-identical to the oracle's
    restatement of Fst.fromMutable (fst.zig:160-224) on random inputs with duplicate arcs."""
    import random
    rng = random.Random(11)
    for _ in range(200):
        spec = random_rhs(rng, max_states=10)
        spec.arcs += spec.arcs[:3]      # exact duplicates
        _, _, img = frozen_pair(L, O, spec)
        assert img == spec.to_oracle(O).freeze().to_bytes()


def test_frozen_queries(L):
    # fst.zig:295-379: freeze/query and the ilabel ordering of a state's arcs
    spec = Spec(3, 0, [None, 2.5, 0.0], [(0, 3, 1, 1.0, 1), (0, 1, 2, 0.5, 2), (0, 3, 0, 1.0, 2), (0, 0, 7, 0.0, 1), (1, 1, 1, 0.0, 2)])
    f = spec.to_product(L).freeze()
    assert f.start() == 0 and f.num_states() == 3 and f.num_arcs(0) == 4 and f.num_arcs(2) == 0 and f.num_arcs(7) == 0
    assert f.arcs(0) == [(0, 7, 0.0, 1), (1, 2, 0.5, 2), (3, 0, 1.0, 2), (3, 1, 1.0, 1)]
    assert f.final_weight(1) == 2.5 and f.final_weight(0) == float("inf") and f.final_weight(9) == float("inf")


def _image(num_states, start, states, arcs, magic=0x46535421, version=1, wt=0):
    b = struct.pack("<IHBBIIII", magic, version, wt, 0, num_states, len(arcs), start, 0)
    for (off, n, fw) in states:
        b += struct.pack("<IId", off, n, fw)
    for (il, ol, w, nx) in arcs:
        b += struct.pack("<IIdII", il, ol, w, nx, 0)
    return b


def test_load_validation_matches_reference_rules(L, O, tmp_path):
    # fst.zig:227-273 and its tests :420-492
    good = _image(2, 0, [(0, 2, float("inf")), (2, 0, 0.0)], [(1, 1, 0.0, 1), (2, 2, 0.0, 1)])
    bad = {
        "magic": _image(2, 0, [(0, 2, 0.0), (2, 0, 0.0)], [(1, 1, 0.0, 1), (2, 2, 0.0, 1)], magic=0x12345678),
        "version": _image(2, 0, [(0, 2, 0.0), (2, 0, 0.0)], [(1, 1, 0.0, 1), (2, 2, 0.0, 1)], version=2),
        "weight_type": _image(2, 0, [(0, 2, 0.0), (2, 0, 0.0)], [(1, 1, 0.0, 1), (2, 2, 0.0, 1)], wt=1),
        "truncated": good[:-8],
        "start_oob": _image(2, 5, [(0, 2, 0.0), (2, 0, 0.0)], [(1, 1, 0.0, 1), (2, 2, 0.0, 1)]),
        "arc_range": _image(2, 0, [(1, 2, 0.0), (2, 0, 0.0)], [(1, 1, 0.0, 1), (2, 2, 0.0, 1)]),
        "arc_offset": _image(2, 0, [(3, 0, 0.0), (2, 0, 0.0)], [(1, 1, 0.0, 1), (2, 2, 0.0, 1)]),
        "target_oob": _image(2, 0, [(0, 2, 0.0), (2, 0, 0.0)], [(1, 1, 0.0, 1), (2, 2, 0.0, 7)]),
        "unsorted": _image(2, 0, [(0, 2, 0.0), (2, 0, 0.0)], [(2, 1, 0.0, 1), (1, 2, 0.0, 1)]),
        "empty_with_start": _image(0, 0, [], []),
        "short": b"FST!",
    }
    p = tmp_path / "x.fst"
    p.write_bytes(good)
    f = L.Fst.load(str(p))
    assert f.num_states() == 2 and f.arcs(0) == [(1, 1, 0.0, 1), (2, 2, 0.0, 1)]
    assert O.Frozen.from_bytes(good).num_states() == 2
    q = tmp_path / "y.fst"
    assert f.save(str(q)) == L.FST_OK and q.read_bytes() == good          # io/binary.zig:9-13 raw dump
    for name, img in bad.items():
        p.write_bytes(img)
        assert L.lib().fst_load(str(p).encode()) == L.FST_INVALID_HANDLE, name
        with pytest.raises(ValueError):
            O.Frozen.from_bytes(img)
    assert L.lib().fst_load(b"/nonexistent/file") == L.FST_INVALID_HANDLE
    assert L.lib().fst_load(None) == L.FST_INVALID_HANDLE
    ok_empty = _image(0, 0xFFFFFFFF, [], [])
    p.write_bytes(ok_empty)
    assert L.Fst.load(str(p)).num_states() == 0
    assert f.save("/nonexistent/dir/x") == L.FST_IO_ERROR


def test_string_helpers(L, O):
    # string.zig:101-183
    for s in (b"hello", b"", "中".encode(), bytes(range(256))):
        m = L.MutableFst.compile_string(s)
        assert m.num_states() == len(s) + 1 and m.start() == 0
        assert m.final_weight(len(s)) == 0.0
        assert m.print_string() == s and m.print_string(True) == s
        assert O.Mutable.compile_string(s).print_string() == s
        if s:
            assert m.arcs(0) == [(s[0] + 1, s[0] + 1, 0.0, 1)]
    lib = L.lib()
    # not a linear chain / empty fst -> -1 ; buffer too small -> -1 (c-api.zig:1352,1367)
    spec = Spec(2, 0, [None, 0.0], [(0, 1, 1, 0.0, 1), (0, 2, 2, 0.0, 1)])
    assert spec.to_product(L).print_string() is None
    assert L.MutableFst().print_string() is None
    m = L.MutableFst.compile_string(b"abc")
    buf = (C.c_uint8 * 2)()
    assert lib.fst_print_string(m.h, buf, 2) == -1
    assert lib.fst_print_string(m.h, None, 3) == 3
    # epsilons are skipped on the chosen tape (string.zig:87-91)
    spec = Spec(3, 0, [None, None, 0.0], [(0, 98, 0, 0.0, 1), (1, 0, 99, 0.0, 2)])
    t = spec.to_product(L)
    assert t.print_string() == b"a" and t.print_string(True) == b"b"
    assert lib.fst_compile_string(None, 0) == L.FST_INVALID_HANDLE


def test_search_calls_fail_loudly_without_a_gpu(L, capfd):
    if L.device_count() > 0:
        pytest.skip("a CUDA device is present")
    f = L.MutableFst.compile_string(b"ab").freeze()
    a = L.MutableFst.compile_string(b"ab")
    assert L.compose_frozen_shortest_path(a, f, 1) is None
    assert "no CPU fallback" in capfd.readouterr().err
    data, off = L.pack_strings([b"ab"])
    with pytest.raises(RuntimeError):
        L.compose_frozen_shortest_path_batch(f, data, off)
    # n == 0 -> empty FST, no search needed (compose-shortest-path.zig:30-32)
    e = L.compose_frozen_shortest_path(a, f, 0)
    assert e is not None and e.num_states() == 0
    out = C.POINTER(L._BatchResult)()
    assert L.lib().fst_compose_frozen_shortest_path_batch(L.FST_INVALID_HANDLE, None, None, 0, C.byref(out)) == L.FST_INVALID_ARG
    # the other batched entries: no device -> FST_INVALID_STATE (3), loudly; bad flags -> FST_INVALID_ARG (2) first
    with pytest.raises(RuntimeError, match="FstError 3"):
        L.compose_frozen_then_shortest_path_batch(f, data, off)
    with pytest.raises(RuntimeError, match="FstError 3"):
        L.compose_frozen_shortest_path_batch_multi(f, data, off)
    assert "no CPU fallback" in capfd.readouterr().err
    with pytest.raises(RuntimeError, match="FstError 2"):
        L.compose_frozen_shortest_path_batch(f, data, off, flags=8)
    with pytest.raises(RuntimeError, match="FstError 2"):
        L.compose_frozen_shortest_path_batch_multi(f, data, off, flags=8)
    assert L.lib().fst_b200_last_path_required() == 0


def test_freed_frozen_handle_is_invalid(L, tmp_path):
    f = L.MutableFst.compile_string(b"xy").freeze()
    h = f.h
    L.lib().fst_free(h)
    f.h = L.FST_INVALID_HANDLE
    assert L.lib().fst_num_states(h) == 0
    assert L.lib().fst_save(h, str(tmp_path / "z").encode()) == L.FST_INVALID_ARG


def test_trace_line_format(L):
    # c-api.zig:74-103 — same stderr line format under LIBFST_TRACE_COMPOSE
    import subprocess
    import sys
    code = ("import libfst_b200 as L; L.load(); a=L.MutableFst.compile_string(b'a'); "
            "L.lib().fst_compose_frozen_shortest_path(a.h, 12345, 1)")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT,
                       env=dict(os.environ, LIBFST_TRACE_COMPOSE="1"))
    assert re.search(r"\[libfst\] sp_invalid_b op=fst_compose_frozen a=\d+ b=12345 in_states=2 in_arcs=1 "
                     r"out_states=0 out_arcs=0 elapsed_us=\d+", r.stderr), r.stderr


def test_configure_validates_its_arguments(L):
    """fst_b200_configure: engines 0..7, lanes 0/4/8/16/32, semantics 0/1; anything else is FST_INVALID_ARG and leaves
    the previous configuration in place."""
    try:
        for engine in range(8):
            L.configure(engine=engine)
        for lanes in (0, 4, 8, 16, 32):
            L.configure(lanes_per_string=lanes)
        L.configure(semantics=L.EAGER)
        for bad in (dict(engine=8), dict(lanes_per_string=5), dict(lanes_per_string=64), dict(semantics=2)):
            with pytest.raises(ValueError):
                L.configure(**bad)
    finally:
        L.configure()
