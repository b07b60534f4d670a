"""libfst_b200 — host-side mirror of the reference C ABI for the hot path.

Thin ctypes binding over ``libfst_b200.so`` (CUDA, sm_100a).  Function names,
argument meaning and error behaviour follow the reference's ``include/fst.h``
(see ``include/libfst_b200.h`` for the per-function citations); the batched entry
points are new.  There is no CPU fallback: if the shared library is missing the
import fails, and if no CUDA device is present every search call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import build as _build

FST_OK, FST_OOM, FST_INVALID_ARG, FST_INVALID_STATE, FST_IO_ERROR = 0, 1, 2, 3, 4
FST_NO_STATE = 0xFFFFFFFF
FST_EPSILON = 0
FST_INVALID_HANDLE = 0xFFFFFFFFFFFFFFFF
PATH, NO_PATH, CYCLE, TOO_LARGE, INTERNAL, NOT_BYTES = 0, 1, 2, 3, 4, 5
RESULT_NO_PATHS = 1


class FstArc(C.Structure):
    _fields_ = [("ilabel", C.c_uint32), ("olabel", C.c_uint32), ("weight", C.c_double), ("nextstate", C.c_uint32)]


class _LatticeResult(C.Structure):
    _fields_ = [("n_strings", C.c_uint32), ("status", C.POINTER(C.c_int32)), ("state_offsets", C.POINTER(C.c_uint64)),
                ("arc_offsets", C.POINTER(C.c_uint64)), ("arc_begin", C.POINTER(C.c_uint32)), ("final_weights", C.POINTER(C.c_double)),
                ("ilabels", C.POINTER(C.c_uint32)), ("olabels", C.POINTER(C.c_uint32)), ("weights", C.POINTER(C.c_double)),
                ("nextstates", C.POINTER(C.c_uint32)), ("device_ms", C.c_double), ("launches", C.c_uint32)]


class _BatchResult(C.Structure):
    _fields_ = [("n_strings", C.c_uint32), ("status", C.POINTER(C.c_int32)), ("path_offsets", C.POINTER(C.c_uint64)),
                ("ilabels", C.POINTER(C.c_uint32)), ("olabels", C.POINTER(C.c_uint32)), ("weights", C.POINTER(C.c_double)),
                ("final_weights", C.POINTER(C.c_double)), ("n_tuples", C.POINTER(C.c_uint32)),
                ("out_offsets", C.POINTER(C.c_uint64)), ("out_bytes", C.POINTER(C.c_uint8)),
                ("device_ms", C.c_double), ("total_tuples", C.c_uint64), ("total_relax", C.c_uint64),
                ("launches", C.c_uint32), ("passes", C.c_uint32)]


class _MultiResult(C.Structure):
    _fields_ = [("n_strings", C.c_uint32), ("n_chunks", C.c_uint32), ("chunk_first", C.POINTER(C.c_uint64)),
                ("chunks", C.POINTER(C.POINTER(_BatchResult))), ("chunk_device", C.POINTER(C.c_int32)), ("n_devices", C.c_uint32),
                ("wall_ms", C.c_double), ("device_ms", C.c_double), ("total_tuples", C.c_uint64), ("total_relax", C.c_uint64),
                ("launches", C.c_uint32)]


class DeviceOut(C.Structure):
    _fields_ = [("d_status", C.c_void_p), ("d_path_offsets", C.c_void_p), ("d_ilabels", C.c_void_p),
                ("d_olabels", C.c_void_p), ("d_weights", C.c_void_p), ("d_final_weights", C.c_void_p),
                ("d_n_tuples", C.c_void_p), ("path_capacity", C.c_uint64)]


class Config(C.Structure):
    _fields_ = [("workspace_bytes", C.c_uint64), ("lanes_per_string", C.c_uint32), ("tuples_hint", C.c_uint32),
                ("exhaustive", C.c_uint32), ("engine", C.c_uint32), ("semantics", C.c_uint32)]


EXPORTS = {
    # name: (restype, argtypes)
    "fst_mutable_new": (C.c_uint64, []),
    "fst_mutable_clone": (C.c_uint64, [C.c_uint64]),
    "fst_mutable_free": (None, [C.c_uint64]),
    "fst_mutable_add_state": (C.c_uint32, [C.c_uint64]),
    "fst_mutable_set_start": (C.c_int, [C.c_uint64, C.c_uint32]),
    "fst_mutable_set_final": (C.c_int, [C.c_uint64, C.c_uint32, C.c_double]),
    "fst_mutable_add_arc": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_uint32]),
    "fst_mutable_start": (C.c_uint32, [C.c_uint64]),
    "fst_mutable_num_states": (C.c_uint32, [C.c_uint64]),
    "fst_mutable_num_arcs": (C.c_uint32, [C.c_uint64, C.c_uint32]),
    "fst_mutable_final_weight": (C.c_double, [C.c_uint64, C.c_uint32]),
    "fst_mutable_get_arcs": (C.c_uint32, [C.c_uint64, C.c_uint32, C.POINTER(FstArc), C.c_uint32]),
    "fst_freeze": (C.c_uint64, [C.c_uint64]),
    "fst_free": (None, [C.c_uint64]),
    "fst_start": (C.c_uint32, [C.c_uint64]),
    "fst_num_states": (C.c_uint32, [C.c_uint64]),
    "fst_num_arcs": (C.c_uint32, [C.c_uint64, C.c_uint32]),
    "fst_final_weight": (C.c_double, [C.c_uint64, C.c_uint32]),
    "fst_get_arcs": (C.c_uint32, [C.c_uint64, C.c_uint32, C.POINTER(FstArc), C.c_uint32]),
    "fst_load": (C.c_uint64, [C.c_char_p]),
    "fst_save": (C.c_int, [C.c_uint64, C.c_char_p]),
    "fst_compose_frozen_shortest_path": (C.c_uint64, [C.c_uint64, C.c_uint64, C.c_uint32]),
    "fst_compile_string": (C.c_uint64, [C.c_char_p, C.c_uint32]),
    "fst_print_string": (C.c_int32, [C.c_uint64, C.POINTER(C.c_uint8), C.c_uint32]),
    "fst_print_output_string": (C.c_int32, [C.c_uint64, C.POINTER(C.c_uint8), C.c_uint32]),
    "fst_teardown": (None, []),
    "fst_compose_frozen_shortest_path_pipeline": (C.c_int, [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32,
                                                            C.POINTER(C.POINTER(_BatchResult))]),
    "fst_compose_frozen_shortest_path_batch": (C.c_int, [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32,
                                                         C.POINTER(C.POINTER(_BatchResult))]),
    "fst_b200_compose_frozen_then_shortest_path_batch": (C.c_int, [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32,
                                                                   C.POINTER(C.POINTER(_BatchResult))]),
    "fst_b200_last_path_required": (C.c_uint64, []),
    "fst_compose_frozen_shortest_path_batch_multi": (C.c_int, [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                                               C.c_uint32, C.c_uint32, C.POINTER(C.POINTER(_MultiResult))]),
    "fst_compose_frozen_shortest_path_batch_ex": (C.c_int, [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                                            C.POINTER(C.POINTER(_BatchResult))]),
    "fst_b200_multi_free": (None, [C.POINTER(_MultiResult)]),
    "fst_b200_batch_free": (None, [C.POINTER(_BatchResult)]),
    "fst_b200_compose_frozen_lattice_batch": (C.c_int, [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32,
                                                        C.POINTER(C.POINTER(_LatticeResult))]),
    "fst_b200_lattice_free": (None, [C.POINTER(_LatticeResult)]),
    "fst_b200_batch_device": (C.c_int, [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                        C.POINTER(DeviceOut), C.c_void_p]),
    "fst_b200_mutable_add_states": (C.c_int, [C.c_uint64, C.c_uint32]),
    "fst_b200_mutable_set_finals": (C.c_int, [C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]),
    "fst_b200_mutable_add_arcs": (C.c_int, [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fst_b200_configure": (C.c_int, [C.POINTER(Config)]),
    "fst_b200_last_counters": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "fst_b200_last_occupancy": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "fst_b200_device_count": (C.c_int32, []),
    "fst_b200_version": (C.c_char_p, []),
}

_lib = None


def load(rebuild_if_stale: bool = True):
    """Load libfst_b200.so (building it with nvcc if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    so = _build.SO
    if rebuild_if_stale and _build.is_stale():
        try:
            _build.build()
        except Exception:
            if not os.path.exists(so):
                raise
    if not os.path.exists(so):
        raise ImportError("libfst_b200.so is missing and could not be built; there is no CPU fallback")
    L = C.CDLL(so)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(L, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def lib():
    return load()


def device_count() -> int:
    return lib().fst_b200_device_count()


ENGINE_AUTO, ENGINE_WARP, ENGINE_LEAN_HASH, ENGINE_LEAN_DENSE = 0, 1, 2, 3


LAZY, EAGER = 0, 1


def configure(workspace_bytes=0, lanes_per_string=0, tuples_hint=0, exhaustive=0, engine=0, semantics=0):
    cfg = Config(workspace_bytes, lanes_per_string, tuples_hint, exhaustive, engine, semantics)
    rc = lib().fst_b200_configure(C.byref(cfg))
    if rc != FST_OK:
        raise ValueError(f"fst_b200_configure failed: {rc}")


def last_occupancy():
    a, b = C.c_uint32(0), C.c_uint32(0)
    lib().fst_b200_last_occupancy(C.byref(a), C.byref(b))
    return dict(resident=a.value, capacity=b.value)


def last_counters():
    a, b, c = C.c_uint32(0), C.c_uint64(0), C.c_double(0)
    lib().fst_b200_last_counters(C.byref(a), C.byref(b), C.byref(c))
    return dict(launches=a.value, relaxations=b.value, device_ms=c.value)


# ── small object layer used by tests/bench (handles are freed on __del__) ──
class MutableFst:
    def __init__(self, handle=None):
        self.h = lib().fst_mutable_new() if handle is None else handle
        if self.h == FST_INVALID_HANDLE:
            raise MemoryError("fst_mutable_new failed")

    def __del__(self):
        try:
            if getattr(self, "h", FST_INVALID_HANDLE) != FST_INVALID_HANDLE and _lib is not None:
                _lib.fst_mutable_free(self.h)
        except Exception:
            pass
        self.h = FST_INVALID_HANDLE

    def add_state(self): return lib().fst_mutable_add_state(self.h)
    def add_states(self, n): return lib().fst_b200_mutable_add_states(self.h, n)
    def set_start(self, s): return lib().fst_mutable_set_start(self.h, s)

    def set_finals(self, states, weights):
        st = np.ascontiguousarray(states, np.uint32); w = np.ascontiguousarray(weights, np.float64)
        return lib().fst_b200_mutable_set_finals(self.h, len(st), st.ctypes.data, w.ctypes.data)

    def add_arcs(self, src, il, ol, w, nxt):
        src, il, ol, nxt = (np.ascontiguousarray(x, np.uint32) for x in (src, il, ol, nxt))
        w = np.ascontiguousarray(w, np.float64)
        return lib().fst_b200_mutable_add_arcs(self.h, len(src), src.ctypes.data, il.ctypes.data, ol.ctypes.data,
                                                w.ctypes.data, nxt.ctypes.data)

    def set_final(self, s, w=0.0): return lib().fst_mutable_set_final(self.h, s, float(w))
    def add_arc(self, src, il, ol, w, nxt): return lib().fst_mutable_add_arc(self.h, src, il, ol, float(w), nxt)
    def start(self): return lib().fst_mutable_start(self.h)
    def num_states(self): return lib().fst_mutable_num_states(self.h)
    def num_arcs(self, s): return lib().fst_mutable_num_arcs(self.h, s)
    def final_weight(self, s): return lib().fst_mutable_final_weight(self.h, s)

    def arcs(self, s):
        n = self.num_arcs(s)
        buf = (FstArc * max(n, 1))()
        k = lib().fst_mutable_get_arcs(self.h, s, buf, n)
        return [(buf[i].ilabel, buf[i].olabel, buf[i].weight, buf[i].nextstate) for i in range(k)]

    def freeze(self) -> "Fst":
        h = lib().fst_freeze(self.h)
        if h == FST_INVALID_HANDLE:
            raise RuntimeError("fst_freeze failed")
        return Fst(h)

    @staticmethod
    def compile_string(b: bytes) -> "MutableFst":
        h = lib().fst_compile_string(b, len(b))
        if h == FST_INVALID_HANDLE:
            raise RuntimeError("fst_compile_string failed")
        return MutableFst(h)

    def print_string(self, output=False):
        buf = (C.c_uint8 * 65536)()
        fn = lib().fst_print_output_string if output else lib().fst_print_string
        n = fn(self.h, buf, 65536)
        return None if n < 0 else bytes(buf[:n])

    def chain(self):
        """Read a linear result chain: (ilabels, olabels, weights, final_weight) or None if empty."""
        if self.start() == FST_NO_STATE:
            return None
        il, ol, w = [], [], []
        s = self.start()
        while self.num_arcs(s) == 1:
            a = self.arcs(s)[0]
            il.append(a[0]); ol.append(a[1]); w.append(a[2]); s = a[3]
        return (np.array(il, np.uint32), np.array(ol, np.uint32), np.array(w, np.float64), self.final_weight(s))


class Fst:
    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        try:
            if getattr(self, "h", FST_INVALID_HANDLE) != FST_INVALID_HANDLE and _lib is not None:
                _lib.fst_free(self.h)
        except Exception:
            pass
        self.h = FST_INVALID_HANDLE

    @staticmethod
    def load(path: str) -> "Fst":
        h = lib().fst_load(path.encode())
        if h == FST_INVALID_HANDLE:
            raise ValueError(f"fst_load({path!r}) failed")
        return Fst(h)

    @staticmethod
    def from_image(image: bytes) -> "Fst":
        """Load a native binary image held in memory (goes through fst_load)."""
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".libfst.fst", delete=False) as f:
            f.write(image)
            p = f.name
        try:
            return Fst.load(p)
        finally:
            os.unlink(p)

    def save(self, path: str): return lib().fst_save(self.h, path.encode())
    def start(self): return lib().fst_start(self.h)
    def num_states(self): return lib().fst_num_states(self.h)
    def num_arcs(self, s): return lib().fst_num_arcs(self.h, s)
    def final_weight(self, s): return lib().fst_final_weight(self.h, s)

    def arcs(self, s):
        n = self.num_arcs(s)
        buf = (FstArc * max(n, 1))()
        k = lib().fst_get_arcs(self.h, s, buf, n)
        return [(buf[i].ilabel, buf[i].olabel, buf[i].weight, buf[i].nextstate) for i in range(k)]


def compose_frozen_shortest_path(a: MutableFst, b: Fst, n: int = 1):
    """fst_compose_frozen_shortest_path; returns a MutableFst or None for FST_INVALID_HANDLE."""
    h = lib().fst_compose_frozen_shortest_path(a.h, b.h, n)
    return None if h == FST_INVALID_HANDLE else MutableFst(h)


@dataclass
class BatchResult:
    status: np.ndarray
    path_offsets: np.ndarray
    ilabels: np.ndarray
    olabels: np.ndarray
    weights: np.ndarray
    final_weights: np.ndarray
    n_tuples: np.ndarray
    out_offsets: np.ndarray
    out_bytes: np.ndarray
    device_ms: float
    total_tuples: int
    total_relax: int
    launches: int
    passes: int

    def path(self, i):
        a, b = int(self.path_offsets[i]), int(self.path_offsets[i + 1])
        return self.ilabels[a:b], self.olabels[a:b], self.weights[a:b]

    def output(self, i):
        a, b = int(self.out_offsets[i]), int(self.out_offsets[i + 1])
        return bytes(self.out_bytes[a:b])

    def total(self, i):
        """Left-to-right sum of the arc weights plus the final weight (what a caller of the chain computes)."""
        t = 0.0
        for x in self.path(i)[2]:
            t += float(x)
        return t + float(self.final_weights[i])


@dataclass
class BatchSummary:
    """What a batched call returned, without copying its arrays (compose_frozen_shortest_path_batch(copy=False))."""
    n_strings: int
    path_arcs: int
    out_bytes: int
    d2h_bytes: int        # bytes of the pinned host result the call filled
    device_ms: float
    total_tuples: int
    total_relax: int
    launches: int
    passes: int


def pack_strings(strings):
    """list[bytes] -> (uint8 data, uint64 offsets[n+1])."""
    lens = np.fromiter((len(s) for s in strings), np.uint64, len(strings))
    offsets = np.zeros(len(strings) + 1, np.uint64)
    np.cumsum(lens, out=offsets[1:])
    data = np.frombuffer(b"".join(strings), np.uint8).copy() if len(strings) and offsets[-1] else np.zeros(0, np.uint8)
    return data, offsets


def compose_frozen_shortest_path_pipeline(first: Fst, second: Fst, data: np.ndarray, offsets: np.ndarray) -> BatchResult:
    """fst_compose_frozen_shortest_path_pipeline: first then second (tagger then verbalizer) without leaving the device."""
    return compose_frozen_shortest_path_batch(first, data, offsets, second=second)


def compose_frozen_then_shortest_path_batch(b: Fst, data: np.ndarray, offsets: np.ndarray) -> BatchResult:
    """fst_b200_compose_frozen_then_shortest_path_batch: the eager pair per string (config 5), semantics per call."""
    return compose_frozen_shortest_path_batch(b, data, offsets, eager=True)


def compose_frozen_shortest_path_batch(b: Fst, data: np.ndarray, offsets: np.ndarray, copy: bool = True, second: Fst = None,
                                       eager: bool = False, flags: int = 0) -> BatchResult:
    """fst_compose_frozen_shortest_path_batch over host buffers (`second`: the two-stage pipeline entry)."""
    data = np.ascontiguousarray(data, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    n = len(offsets) - 1
    keep = data if data.size else np.zeros(1, np.uint8)
    out = C.POINTER(_BatchResult)()
    if flags:
        rc = lib().fst_compose_frozen_shortest_path_batch_ex(b.h, keep.ctypes.data, offsets.ctypes.data, n, flags, C.byref(out))
    elif eager:
        rc = lib().fst_b200_compose_frozen_then_shortest_path_batch(b.h, keep.ctypes.data, offsets.ctypes.data, n, C.byref(out))
    elif second is None:
        rc = lib().fst_compose_frozen_shortest_path_batch(b.h, keep.ctypes.data, offsets.ctypes.data, n, C.byref(out))
    else:
        rc = lib().fst_compose_frozen_shortest_path_pipeline(b.h, second.h, keep.ctypes.data, offsets.ctypes.data, n, C.byref(out))
    if rc != FST_OK:
        raise RuntimeError(f"fst_compose_frozen_shortest_path_batch failed: FstError {rc}")
    try:
        if not copy:
            # summary only (bench: the timed region is the C-ABI call, not numpy copies of its pinned result)
            r = out.contents
            pt, ot = (int(r.path_offsets[n]), int(r.out_offsets[n])) if n else (0, 0)
            if not r.ilabels:
                pt = 0
            return BatchSummary(n, pt, ot, n * (4 + 8 + 4) + 2 * (n + 1) * 8 + pt * 16 + ot, r.device_ms, r.total_tuples, r.total_relax,
                                r.launches, r.passes)
        return _copy_batch_result(out.contents, n)
    finally:
        lib().fst_b200_batch_free(out)


def _copy_batch_result(r, n) -> BatchResult:
    def arr(ptr, cnt, dt):
        if cnt == 0:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(ptr, shape=(cnt,)).astype(dt, copy=True)
    poff = arr(r.path_offsets, n + 1, np.uint64)
    ooff = arr(r.out_offsets, n + 1, np.uint64)
    total, ototal = int(poff[-1]), int(ooff[-1])
    if not r.ilabels:          # FST_B200_RESULT_NO_PATHS: lengths only
        total = 0
    return BatchResult(arr(r.status, n, np.int32), poff, arr(r.ilabels, total, np.uint32), arr(r.olabels, total, np.uint32),
                       arr(r.weights, total, np.float64), arr(r.final_weights, n, np.float64), arr(r.n_tuples, n, np.uint32),
                       ooff, arr(r.out_bytes, ototal, np.uint8), r.device_ms, r.total_tuples, r.total_relax, r.launches, r.passes)


@dataclass
class MultiResult:
    """fst_compose_frozen_shortest_path_batch_multi: the chunks in input order (each an ordinary BatchResult)."""
    chunk_first: np.ndarray
    chunk_device: np.ndarray
    chunks: list
    n_devices: int
    wall_ms: float
    device_ms: float
    total_tuples: int
    total_relax: int
    launches: int

    def flat(self) -> BatchResult:
        """All chunks concatenated into one BatchResult (a host copy; tests compare this with the 1-GPU result)."""
        if not self.chunks:
            z = np.zeros(1, np.uint64)
            return BatchResult(np.zeros(0, np.int32), z, np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0), np.zeros(0),
                               np.zeros(0, np.uint32), z.copy(), np.zeros(0, np.uint8), 0.0, 0, 0, 0, 0)
        poff, ooff, pb, ob = [np.zeros(1, np.uint64)], [np.zeros(1, np.uint64)], 0, 0
        for c in self.chunks:
            poff.append(c.path_offsets[1:] + np.uint64(pb)); ooff.append(c.out_offsets[1:] + np.uint64(ob))
            pb += int(c.path_offsets[-1]); ob += int(c.out_offsets[-1])
        cat = lambda f: np.concatenate([getattr(c, f) for c in self.chunks])   # noqa: E731
        return BatchResult(cat("status"), np.concatenate(poff), cat("ilabels"), cat("olabels"), cat("weights"), cat("final_weights"),
                           cat("n_tuples"), np.concatenate(ooff), cat("out_bytes"), self.device_ms, self.total_tuples, self.total_relax,
                           self.launches, sum(c.passes for c in self.chunks))


def compose_frozen_shortest_path_batch_multi(b: Fst, data: np.ndarray, offsets: np.ndarray, devices=None, chunks_per_device: int = 0,
                                             copy: bool = True, flags: int = 0):
    """fst_compose_frozen_shortest_path_batch_multi.  copy=False returns only the summary (counters, chunk layout) and frees
    the native result at once (bench: the timed region must not include numpy copies)."""
    data = np.ascontiguousarray(data, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    n = len(offsets) - 1
    keep = data if data.size else np.zeros(1, np.uint8)
    dev = None if devices is None else np.ascontiguousarray(devices, np.int32)
    out = C.POINTER(_MultiResult)()
    rc = lib().fst_compose_frozen_shortest_path_batch_multi(b.h, keep.ctypes.data, offsets.ctypes.data, n,
                                                            None if dev is None else dev.ctypes.data, 0 if dev is None else len(dev),
                                                            chunks_per_device, flags, C.byref(out))
    if rc != FST_OK:
        raise RuntimeError(f"fst_compose_frozen_shortest_path_batch_multi failed: FstError {rc}")
    r = out.contents
    try:
        k = r.n_chunks
        first = np.ctypeslib.as_array(r.chunk_first, shape=(k + 1,)).astype(np.uint64, copy=True)
        cdev = np.ctypeslib.as_array(r.chunk_device, shape=(k,)).astype(np.int32, copy=True) if k else np.zeros(0, np.int32)
        chunks = []
        d2h = 0
        for i in range(k):
            c = r.chunks[i].contents
            cn = int(first[i + 1] - first[i])
            if copy:
                chunks.append(_copy_batch_result(c, cn))
            else:
                pt, ot = (int(c.path_offsets[cn]) if c.ilabels else 0), int(c.out_offsets[cn])
                d2h += cn * (4 + 8 + 4) + 2 * (cn + 1) * 8 + pt * 16 + ot
        m = MultiResult(first, cdev, chunks, r.n_devices, r.wall_ms, r.device_ms, r.total_tuples, r.total_relax, r.launches)
        m.d2h_bytes = d2h
        return m
    finally:
        lib().fst_b200_multi_free(out)


class LatticeBatch:
    """Eager lattices of a batch (fst_b200_compose_frozen_lattice_batch), copied out of the library's buffers."""

    def __init__(self, status, state_offsets, arc_offsets, arc_begin, finals, il, ol, w, nxt, device_ms):
        self.status, self.state_offsets, self.arc_offsets = status, state_offsets, arc_offsets
        self.arc_begin, self.finals, self.ilabels, self.olabels, self.weights, self.nextstates = arc_begin, finals, il, ol, w, nxt
        self.device_ms = device_ms

    def lattice(self, i):
        """(arc_begin[n_states + 1] relative, finals, ilabels, olabels, weights, nextstates) of string i."""
        s0, s1 = int(self.state_offsets[i]), int(self.state_offsets[i + 1])
        a0, a1 = int(self.arc_offsets[i]), int(self.arc_offsets[i + 1])
        ab = np.concatenate([self.arc_begin[s0:s1].astype(np.uint64), np.array([a1 - a0], np.uint64)])
        return ab, self.finals[s0:s1], self.ilabels[a0:a1], self.olabels[a0:a1], self.weights[a0:a1], self.nextstates[a0:a1]


def compose_frozen_lattice_batch(b: Fst, data: np.ndarray, offsets: np.ndarray) -> LatticeBatch:
    data = np.ascontiguousarray(data, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    n = len(offsets) - 1
    keep = data if data.size else np.zeros(1, np.uint8)
    out = C.POINTER(_LatticeResult)()
    rc = lib().fst_b200_compose_frozen_lattice_batch(b.h, keep.ctypes.data, offsets.ctypes.data, n, C.byref(out))
    if rc != FST_OK:
        raise RuntimeError(f"fst_b200_compose_frozen_lattice_batch failed: FstError {rc}")
    r = out.contents
    try:
        def arr(ptr, cnt, dt):
            return np.ctypeslib.as_array(ptr, shape=(cnt,)).astype(dt, copy=True) if cnt else np.zeros(0, dt)
        so, ao = arr(r.state_offsets, n + 1, np.uint64), arr(r.arc_offsets, n + 1, np.uint64)
        S, A = int(so[-1]), int(ao[-1])
        return LatticeBatch(arr(r.status, n, np.int32), so, ao, arr(r.arc_begin, S, np.uint32), arr(r.final_weights, S, np.float64),
                            arr(r.ilabels, A, np.uint32), arr(r.olabels, A, np.uint32), arr(r.weights, A, np.float64),
                            arr(r.nextstates, A, np.uint32), r.device_ms)
    finally:
        lib().fst_b200_lattice_free(out)


def teardown():
    lib().fst_teardown()
