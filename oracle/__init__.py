"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of the C++ restatement in ``fst_oracle.hpp`` (each function there
cites the reference file:line it follows).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product package ``libfst_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

STATUS_OK, STATUS_EMPTY, STATUS_UNSUPPORTED_N, STATUS_BACKTRACK_CYCLE = 0, 1, 2, 3
KIND_PLAIN, KIND_EPS_DENSE, KIND_AMBIGUOUS = 0, 1, 2
NO_STATE = 0xFFFFFFFF


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (seconds).  Idempotent."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "fst_oracle.hpp", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


class _Info(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_arcs", C.c_uint32), ("final_weight", C.c_double),
                ("total", C.c_double), ("tuples", C.c_uint64), ("relax_calls", C.c_uint64),
                ("pushes", C.c_uint64), ("retakes", C.c_uint64), ("pops", C.c_uint64),
                ("final_s1", C.c_uint32), ("final_s2", C.c_uint32), ("final_filter", C.c_uint32),
                ("_pad", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    try:
        L = C.CDLL(_SO)
    except OSError:
        build(force=True)
        L = C.CDLL(_SO)
    vp, u32, u64, dbl = C.c_void_p, C.c_uint32, C.c_uint64, C.c_double
    pu8, pu32, pu64, pd, pi32 = (C.POINTER(C.c_uint8), C.POINTER(u32), C.POINTER(u64), C.POINTER(dbl),
                                 C.POINTER(C.c_int32))
    sig = {
        "orc_mutable_new": (vp, []),
        "orc_mutable_free": (None, [vp]),
        "orc_mutable_add_state": (u32, [vp]),
        "orc_mutable_set_start": (None, [vp, u32]),
        "orc_mutable_set_final": (None, [vp, u32, dbl]),
        "orc_mutable_add_arc": (None, [vp, u32, u32, u32, dbl, u32]),
        "orc_mutable_from_arrays": (vp, [u32, u32, pd, u32, pu32, pu32, pu32, pd, pu32]),
        "orc_compile_string": (vp, [pu8, u32]),
        "orc_compile_string_transducer": (vp, [pu8, u32, pu8, u32]),
        "orc_mutable_num_states": (u32, [vp]),
        "orc_freeze": (vp, [vp]),
        "orc_fst_from_bytes": (vp, [pu8, u64]),
        "orc_fst_num_bytes": (u64, [vp]),
        "orc_fst_copy_bytes": (None, [vp, pu8]),
        "orc_fst_free": (None, [vp]),
        "orc_fst_num_states": (u32, [vp]),
        "orc_fst_num_arcs_total": (u32, [vp]),
        "orc_gen_frozen": (vp, [C.c_int, u32, u32]),
        "orc_csp_mutable": (None, [vp, vp, u32, u32, pu32, pu32, pd, C.POINTER(_Info)]),
        "orc_csp_bytes": (None, [vp, pu8, u32, u32, pu32, pu32, pd, C.POINTER(_Info)]),
        "orc_eager_mutable": (None, [vp, vp, u32, u32, pu32, pu32, pd, C.POINTER(_Info), pu64, pu64]),
        "orc_csp_batch_bytes": (dbl, [vp, pu8, pu64, u32, u32, u32, pu32, pu32, pd, pi32, pu32, pd, pd, pu64, pu64]),
        "orc_eager_batch_bytes": (dbl, [vp, pu8, pu64, u32, u32, u32, pu32, pu32, pd, pi32, pu32, pd, pd, pu64, pu64]),
        "orc_print_string": (C.c_int32, [vp, C.c_int, pu8, u32]),
        "orc_compose_bytes": (vp, [vp, pu8, u32]),
        "orc_mutable_total_arcs": (C.c_uint64, [vp]),
        "orc_mutable_start": (u32, [vp]),
        "orc_mutable_dump": (None, [vp, pu64, pd, pu32, pu32, pd, pu32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


@dataclass
class Path:
    status: int
    ilabels: np.ndarray
    olabels: np.ndarray
    weights: np.ndarray
    final_weight: float
    total: float
    tuples: int = 0
    relax_calls: int = 0
    pushes: int = 0
    retakes: int = 0
    pops: int = 0
    final_tuple: tuple = field(default=(NO_STATE, NO_STATE, 0))

    def signature(self) -> str:
        """SURVEY.md App. B ``sha16``: 'il,ol,w' items joined by ';', w like repr(float)."""
        s = ";".join(f"{int(i)},{int(o)},{float(w)!r}" for i, o, w in zip(self.ilabels, self.olabels, self.weights))
        return hashlib.sha256(s.encode()).hexdigest()[:16]

    def output_bytes(self):
        """string.zig:64-97 on the output tape of the result chain."""
        if self.status != STATUS_OK:
            return None
        ol = self.olabels[self.olabels != 0]
        if (ol > 256).any():
            return None
        return bytes((ol - 1).astype(np.uint8))


class Mutable:
    def __init__(self, ptr=None):
        self.ptr = ptr if ptr is not None else lib().orc_mutable_new()

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().orc_mutable_free(self.ptr)
            self.ptr = None

    def add_state(self): return lib().orc_mutable_add_state(self.ptr)
    def add_states(self, n):
        for _ in range(n): self.add_state()
    def set_start(self, s): lib().orc_mutable_set_start(self.ptr, s)
    def set_final(self, s, w=0.0): lib().orc_mutable_set_final(self.ptr, s, float(w))
    def add_arc(self, src, il, ol, w, nxt): lib().orc_mutable_add_arc(self.ptr, src, il, ol, float(w), nxt)
    def num_states(self): return lib().orc_mutable_num_states(self.ptr)

    @staticmethod
    def from_arrays(num_states, start, finals, src, il, ol, w, nxt):
        finals = np.ascontiguousarray(finals, np.float64)
        src, il, ol, nxt = (np.ascontiguousarray(x, np.uint32) for x in (src, il, ol, nxt))
        w = np.ascontiguousarray(w, np.float64)
        return Mutable(lib().orc_mutable_from_arrays(num_states, start, _p(finals, C.c_double), len(src),
                                                     _p(src, C.c_uint32), _p(il, C.c_uint32), _p(ol, C.c_uint32),
                                                     _p(w, C.c_double), _p(nxt, C.c_uint32)))

    @staticmethod
    def compile_string(b: bytes):
        a = np.frombuffer(b, np.uint8) if len(b) else np.zeros(1, np.uint8)
        return Mutable(lib().orc_compile_string(_p(a, C.c_uint8), len(b)))

    @staticmethod
    def compile_string_transducer(i: bytes, o: bytes):
        a = np.frombuffer(i, np.uint8) if len(i) else np.zeros(1, np.uint8)
        b = np.frombuffer(o, np.uint8) if len(o) else np.zeros(1, np.uint8)
        return Mutable(lib().orc_compile_string_transducer(_p(a, C.c_uint8), len(i), _p(b, C.c_uint8), len(o)))

    def freeze(self) -> "Frozen":
        return Frozen(lib().orc_freeze(self.ptr))

    def dump(self):
        """(start, arc_begin[n+1], finals[n], ilabel, olabel, weight, nextstate) in stored order."""
        n, a = self.num_states(), lib().orc_mutable_total_arcs(self.ptr)
        ab, fin = np.zeros(n + 1, np.uint64), np.zeros(max(n, 1), np.float64)
        il, ol, nx, w = np.zeros(max(a, 1), np.uint32), np.zeros(max(a, 1), np.uint32), np.zeros(max(a, 1), np.uint32), np.zeros(max(a, 1), np.float64)
        lib().orc_mutable_dump(self.ptr, _p(ab, C.c_uint64), _p(fin, C.c_double), _p(il, C.c_uint32), _p(ol, C.c_uint32), _p(w, C.c_double), _p(nx, C.c_uint32))
        return lib().orc_mutable_start(self.ptr), ab, fin[:n], il[:a], ol[:a], w[:a], nx[:a]

    def print_string(self, output_tape=False):
        buf = np.zeros(1 << 16, np.uint8)
        n = lib().orc_print_string(self.ptr, int(output_tape), _p(buf, C.c_uint8), len(buf))
        return None if n < 0 else bytes(buf[:n])


class Frozen:
    def __init__(self, ptr):
        if not ptr:
            raise ValueError("invalid frozen image")
        self.ptr = ptr

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().orc_fst_free(self.ptr)
            self.ptr = None

    @staticmethod
    def generate(kind: int, T: int, B: int) -> "Frozen":
        return Frozen(lib().orc_gen_frozen(kind, T, B))

    @staticmethod
    def from_bytes(b: bytes) -> "Frozen":
        a = np.frombuffer(b, np.uint8)
        return Frozen(lib().orc_fst_from_bytes(_p(a, C.c_uint8), len(b)))

    def to_bytes(self) -> bytes:
        n = lib().orc_fst_num_bytes(self.ptr)
        a = np.zeros(n, np.uint8)
        lib().orc_fst_copy_bytes(self.ptr, _p(a, C.c_uint8))
        return a.tobytes()

    def num_states(self): return lib().orc_fst_num_states(self.ptr)
    def num_arcs(self): return lib().orc_fst_num_arcs_total(self.ptr)


def _mk_path(info, il, ol, w):
    n = min(info.n_arcs, len(il))
    return Path(info.status, il[:n].copy(), ol[:n].copy(), w[:n].copy(), info.final_weight, info.total,
                info.tuples, info.relax_calls, info.pushes, info.retakes, info.pops,
                (info.final_s1, info.final_s2, info.final_filter))


def csp_mutable(lhs: Mutable, fst: Frozen, n: int = 1, cap: int = 1 << 16) -> Path:
    il, ol, w = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.float64)
    info = _Info()
    lib().orc_csp_mutable(lhs.ptr, fst.ptr, n, cap, _p(il, C.c_uint32), _p(ol, C.c_uint32), _p(w, C.c_double), C.byref(info))
    return _mk_path(info, il, ol, w)


def csp_bytes(fst: Frozen, s: bytes, cap: int = 1 << 16) -> Path:
    a = np.frombuffer(s, np.uint8) if len(s) else np.zeros(1, np.uint8)
    il, ol, w = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.float64)
    info = _Info()
    lib().orc_csp_bytes(fst.ptr, _p(a, C.c_uint8), len(s), cap, _p(il, C.c_uint32), _p(ol, C.c_uint32), _p(w, C.c_double), C.byref(info))
    return _mk_path(info, il, ol, w)


def eager_mutable(lhs: Mutable, fst: Frozen, n: int = 1, cap: int = 1 << 16):
    il, ol, w = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.float64)
    info = _Info()
    ls, la = C.c_uint64(0), C.c_uint64(0)
    lib().orc_eager_mutable(lhs.ptr, fst.ptr, n, cap, _p(il, C.c_uint32), _p(ol, C.c_uint32), _p(w, C.c_double),
                            C.byref(info), C.byref(ls), C.byref(la))
    return _mk_path(info, il, ol, w), ls.value, la.value


def compose_bytes(fst: Frozen, s: bytes) -> Mutable:
    """Eager lattice compose(compile_string(s), fst) (compose.zig:29-198)."""
    a = np.frombuffer(s, np.uint8) if len(s) else np.zeros(1, np.uint8)
    return Mutable(lib().orc_compose_bytes(fst.ptr, _p(a, C.c_uint8), len(s)))


def csp_batch_bytes(fst: Frozen, data: np.ndarray, offsets: np.ndarray, n_threads: int = 1, cap: int = 0, eager: bool = False):
    """Thread-pool batch (CPU baseline).  Returns dict with seconds, per-string arrays, work sums.
    eager=True runs compose() + shortestPath() per string instead of composeShortestPath()."""
    data = np.ascontiguousarray(data, np.uint8)
    if data.size == 0:
        data = np.zeros(1, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    n = len(offsets) - 1
    status, lens = np.zeros(n, np.int32), np.zeros(n, np.uint32)
    finals, totals = np.zeros(n, np.float64), np.zeros(n, np.float64)
    il = ol = w = None
    if cap:
        il, ol, w = np.zeros(n * cap, np.uint32), np.zeros(n * cap, np.uint32), np.zeros(n * cap, np.float64)
    st, sr = C.c_uint64(0), C.c_uint64(0)
    secs = (lib().orc_eager_batch_bytes if eager else lib().orc_csp_batch_bytes)(fst.ptr, _p(data, C.c_uint8), _p(offsets, C.c_uint64), n, n_threads, cap,
                                     _p(il, C.c_uint32), _p(ol, C.c_uint32), _p(w, C.c_double),
                                     _p(status, C.c_int32), _p(lens, C.c_uint32), _p(finals, C.c_double),
                                     _p(totals, C.c_double), C.byref(st), C.byref(sr))
    return dict(seconds=secs, status=status, lens=lens, finals=finals, totals=totals, ilabels=il, olabels=ol,
                weights=w, cap=cap, tuples=st.value, relax_calls=sr.value)
