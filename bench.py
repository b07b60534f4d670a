#!/usr/bin/env python3
"""bench.py — batched compose_shortest_path throughput on B200.

    python bench.py --gpus N --steps K --warmup W              # headline (BASELINE config 2), weak scaling under torchrun
    python bench.py --config 1|2|3:L|3a:L|4|5|mixed|mixed-amb|plain   # one BASELINE config (3:L = eps-dense len L, 3a:L = ambiguous len L)
    python bench.py --matrix [--out FILE]                       # every config, one JSON line each (subprocess per config)
    python bench.py --scaling strong --gpus N [--total T]       # ONE batch split over N GPUs by the product's multi-GPU entry
    python bench.py --impl reference ...                        # CPU arm: oracle port of the reference on all host cores

One "step" = one pass of the hot path over one batch of synthetic strings.  Default workload = BASELINE.json
configs[1]: scenario compose_frozen_lazy_shortest_path_epsilon_dense, len 96, transducer-len 4096, branches 12; the
literal 1 M-string batch is processed as consecutive steps of `--batch` strings per GPU (throughput is per string; the
batch actually run is in the JSON).  Under torchrun (weak scaling) every rank searches its own `--batch` strings
against its own replica of the transducer; there is no collective on the data path.

`value` : strings/s with inputs and outputs resident in HBM (fst_b200_batch_device), CUDA events on the launch stream.
`e2e`   : strings/s through fst_compose_frozen_shortest_path_batch with HOST buffers (H2D of the strings and D2H of
          the paths inside the timed region).  Strong scaling: both through fst_compose_frozen_shortest_path_batch_multi
          (`value` from the largest per-device device time, `e2e` from the wall time of the call).
Every timed batch is verified: strings that are equal must give equal results (checked on the device over the WHOLE
batch) and one string of every class — or 64 strings of a batch of distinct strings — is compared bit for bit with
the CPU oracle; `work_per_string.checked_vs_oracle` is the number of strings covered.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"epsilon_dense": 1, "ambiguous": 2, "plain": 0, "wetext": -1}   # name -> oracle generator kind (CPU arm only)
SCENARIO = {"epsilon_dense": "compose_frozen_lazy_shortest_path_epsilon_dense",
            "ambiguous": "compose_frozen_lazy_shortest_path_ambiguous",
            "plain": "compose_frozen_lazy_shortest_path",
            "wetext": "synthetic WeText-style tagger (SURVEY 8d config 4)"}
MIXED_LENS = [11, 19, 33, 64, 96, 128, 160, 192, 224, 251]   # bench/run_issue1_profile_bench.py:24-25
DEFAULT_BATCH = {"epsilon_dense": 9472, "ambiguous": 65536, "plain": 1 << 20, "wetext": 1 << 18}
# --matrix: BASELINE configs 1-5 at their literal shapes (3 = the issue-#1 length sweep on both bench transducers)
MATRIX = (["1", "2"] + [f"3:{l}" for l in (11, 33, 64, 128, 192, 251)] + [f"3a:{l}" for l in (11, 33, 160, 251)] +
          ["mixed", "mixed-amb", "plain", "4", "5"])


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="", help="BASELINE config: 1, 2, 3:L, 3a:L, 4, 5, mixed, mixed-amb, plain (sets the flags below)")
    ap.add_argument("--matrix", action="store_true", help="run every config in a subprocess, one JSON line each")
    ap.add_argument("--out", default="", help="--matrix: also append the lines to this file")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total", type=int, default=0, help="--scaling strong: strings of the one batch (0 = workload default)")
    ap.add_argument("--distinct", type=int, default=1 << 20, help="wetext, --scaling strong: distinct strings generated (tiled up to --total)")
    ap.add_argument("--chunks-per-device", type=int, default=0)
    ap.add_argument("--bytes-only", action="store_true",
                    help="--scaling strong: FST_B200_RESULT_NO_PATHS (output strings, statuses, path lengths; no per-arc arrays in the D2H)")
    ap.add_argument("--workload", default="epsilon_dense", choices=sorted(WORKLOADS))
    ap.add_argument("--len", type=int, default=96)
    ap.add_argument("--transducer-len", type=int, default=4096)
    ap.add_argument("--branches", type=int, default=12)
    ap.add_argument("--batch", type=int, default=0, help="strings per GPU per step (0 = workload default)")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--exhaustive", type=int, default=0)
    ap.add_argument("--dict", type=int, default=110000, help="wetext workload: dictionary entries (110000 ~ 1 M arcs)")
    ap.add_argument("--tuples-hint", type=int, default=0, help="expected tuples per string (0 = adaptive: learnt in warm-up)")
    ap.add_argument("--semantics", default="lazy", choices=["lazy", "eager"],
                    help="lazy = fst_compose_frozen_shortest_path (headline); eager = compose then shortest_path (config 5)")
    ap.add_argument("--engine", type=int, default=0, help="0 auto, 1 general warp kernel, 2 lean+hash, 3 lean+dense")
    ap.add_argument("--mixed", action="store_true",
                    help="lengths drawn uniformly from the issue #1 profile list {11..251} (seed 1) instead of --len: load balance")
    ap.add_argument("--latency", action="store_true", help="also time the single-call drop-in (one string per call) next to the CPU port")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="strings in the CPU baseline sample (0 = auto)")
    args = ap.parse_args(argv)
    apply_config(args)
    return args


def apply_config(args):
    c = args.config
    if not c:
        return
    if c == "1":
        args.workload, args.len, args.latency = "ambiguous", 96, True
    elif c == "2":
        args.workload, args.len = "epsilon_dense", 96
    elif c.startswith("3a:"):
        args.workload, args.len = "ambiguous", int(c[3:])
    elif c.startswith("3:"):
        args.workload, args.len = "epsilon_dense", int(c[2:])
    elif c == "4":
        args.workload = "wetext"
    elif c == "5":
        args.workload, args.len, args.semantics = "ambiguous", 251, "eager"
    elif c == "mixed":
        args.workload, args.mixed = "epsilon_dense", True
    elif c == "mixed-amb":
        args.workload, args.mixed = "ambiguous", True
    elif c == "plain":
        args.workload, args.len = "plain", 96
    else:
        raise SystemExit(f"bench.py: unknown --config {c!r}")


def input_string(workload: str, length: int, branches: int) -> bytes:
    if workload == "plain":
        return bytes(i % max(1, branches) for i in range(length))
    return bytes(length)


def workload_strings(args, batch, seed, sources=None):
    """(uint8 data, uint64 offsets, max_len) of `batch` input strings of the workload (numpy only)."""
    from libfst_b200 import synth
    if args.workload == "wetext":
        distinct = min(batch, args.distinct)
        data, offsets = synth.wetext_packed(sources, distinct, seed=seed)
        if distinct < batch:   # a larger batch repeats the distinct strings (every copy is searched independently)
            reps = -(-batch // distinct)
            lens = np.tile(np.diff(offsets.astype(np.int64)), reps)[:batch]
            data = np.tile(data, reps)[:int(lens.sum())]
            offsets = np.zeros(batch + 1, np.uint64); np.cumsum(lens.astype(np.uint64), out=offsets[1:])
    elif getattr(args, "mixed", False):
        lens = np.random.default_rng(seed).choice(np.array(MIXED_LENS, np.uint64), batch)
        offsets = np.zeros(batch + 1, np.uint64); np.cumsum(lens, out=offsets[1:])
        full = np.frombuffer(input_string(args.workload, max(MIXED_LENS), args.branches), np.uint8)
        data = np.concatenate([full[:int(n)] for n in lens]) if batch else np.zeros(0, np.uint8)
    else:
        s = input_string(args.workload, args.len, args.branches)
        data = np.frombuffer(s * batch, np.uint8) if len(s) else np.zeros(0, np.uint8)
        offsets = np.arange(batch + 1, dtype=np.uint64) * len(s)
    max_len = int(np.diff(offsets.astype(np.int64)).max()) if batch else 0
    return data, offsets, max_len


def make_workload(args, batch, seed):
    """Product arm: (Fst built through the C ABI, data, offsets, max_len, oracle_loader, sources).  `oracle_loader()`
    (checker / cpu_baseline leg only) gives the oracle's copy of the same frozen image."""
    from libfst_b200 import synth
    sources = None
    if args.workload == "wetext":
        m, sources = synth.wetext_style(args.dict)
        fst = m.freeze()
    else:
        fst = synth.TRANSDUCERS[args.workload](args.transducer_len, args.branches).freeze()
    data, offsets, max_len = workload_strings(args, batch, seed, sources)

    def oracle_loader():
        import oracle
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".fst", delete=False) as t:
            path = t.name
        try:
            assert fst.save(path) == 0
            return oracle.Frozen.from_bytes(open(path, "rb").read())
        finally:
            os.unlink(path)
    return fst, data, offsets, max_len, oracle_loader, sources


def make_reference_workload(args):
    """CPU arm: the transducer built with the oracle only (nothing of the product library is loaded)."""
    import oracle
    from libfst_b200 import synth     # numpy generators only; does not load libfst_b200.so
    if args.workload == "wetext":
        n_states, src, il, ol, w, nxt, sources = synth.wetext_arrays(args.dict)
        finals = np.full(n_states, np.inf); finals[0] = 0.0
        return oracle.Mutable.from_arrays(n_states, 0, finals, src, il, ol, w, nxt).freeze(), sources
    return oracle.Frozen.generate(WORKLOADS[args.workload], args.transducer_len, args.branches), None


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks/throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_sample_size(args, cores, one, batch=None, seconds=10.0):
    """Strings of the bounded CPU sample: about `seconds` of work on all cores, a multiple of the core count (no thread
    ends its share with a partial tail)."""
    if args.cpu_sample:
        return args.cpu_sample
    n = int(max(cores, min(cores * 4096, cores * max(1.0, seconds / one))))
    if batch is not None:
        n = min(n, batch)
    return max(cores, (n // cores) * cores) if n >= cores else n


def run_reference(args):
    """CPU arm: the oracle port of the reference's composeShortestPath on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    cores = os.cpu_count() or 1
    f, sources = make_reference_workload(args)
    probe_n = 64 if args.workload == "wetext" else 1
    pdata, poff, _ = workload_strings(args, probe_n, 1, sources)
    eager = args.semantics == "eager"
    t0 = time.time(); oracle.csp_batch_bytes(f, pdata, poff, n_threads=1, eager=eager); one = max((time.time() - t0) / probe_n, 1e-7)
    # bounded sample per step: a few seconds of work on all cores
    sample = cpu_sample_size(args, cores, one, seconds=5.0)
    data, offsets, _ = workload_strings(args, sample, 1, sources)
    R1 = 0.0
    for _ in range(args.warmup):
        oracle.csp_batch_bytes(f, data, offsets, n_threads=cores, eager=eager)
    secs = 0.0
    for _ in range(args.steps):
        r = oracle.csp_batch_bytes(f, data, offsets, n_threads=cores, eager=eager)
        secs += r["seconds"]; R1 = r["relax_calls"] / sample
    ms = secs / args.steps * 1e3
    v = sample / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "strings/sec batched compose_shortest_path", "value": v, "unit": "strings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, sample, 0), cache="n/a (CPU arm)"),
        "composed_arcs_per_sec": v * R1,
        "cpu_baseline": {"value": v, "unit": "strings/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} strings/step of the same workload, {cores} threads, C++ restatement of the reference "
                                   f"(zig toolchain absent; oracle/fst_oracle.hpp)"},
        "e2e": {"value": v, "unit": "strings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, batch, state_bytes, resident=None):
    desc = (f"{SCENARIO[args.workload]} dict={args.dict} len=U[11,251] (70% dictionary words, 30% printable bytes)"
            if args.workload == "wetext" else
            f"{SCENARIO[args.workload]} len={'mixed U{11,19,33,64,96,128,160,192,224,251} seed 1' if getattr(args, 'mixed', False) else args.len} transducer_len={args.transducer_len} branches={args.branches}")
    if args.semantics == "eager":
        desc = desc.replace("compose_frozen_lazy_shortest_path", "compose_frozen") + " + shortest_path (eager lattice, config 5)"
    return {"workload": desc, "baseline_config": args.config or ("2" if args.workload == "epsilon_dense" and args.len == 96 and not args.mixed else None),
            "semantics": args.semantics,
            "batch_per_gpu_per_step": batch, "resident_strings_per_gpu": resident, "literal_batch": 10000000 if args.workload == "wetext" else 1000000,
            "cache": (f"per-step search state ~{state_bytes / 2**30:.1f} GiB in HBM >> 126 MB L2, rewritten by every string; "
                      f"no L2 flush needed") if state_bytes > (1 << 30) else
                     "search state fits L2: a buffer larger than L2 is written between timed steps",
            "lanes_per_string": args.lanes or "auto", "exhaustive": args.exhaustive, "engine": args.engine or "auto"}


def transducer_size(fst):
    """States and arcs of the frozen transducer, read back through the C ABI (fst_num_states / fst_num_arcs)."""
    n = int(fst.num_states())
    return {"states": n, "arcs": int(sum(fst.num_arcs(s) for s in range(n)))}


def traffic_key(args):
    if args.workload == "wetext":
        return f"wetext:{args.dict}"
    return (f"{args.workload}:{'mixed' if args.mixed else args.len}:{args.transducer_len}:{args.branches}" +
            (":eager" if args.semantics == "eager" else ""))


def measured_traffic(args):
    """DRAM / L2 bytes per string of the search kernel from the committed ncu capture of this workload
    (profiles/traffic.json, written by scripts/traffic.py from `ncu --metrics dram__bytes_*`): NOT measured in this run."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(traffic_key(args))
        return t if isinstance(t, dict) else None
    except Exception:
        return None


_JSON_FD = None


def emit(line: dict):
    """The one JSON line of the run, on the process's ORIGINAL stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def oracle_path(oracle, f, s: bytes, eager: bool):
    if eager:
        p1, ls, la = oracle.eager_mutable(oracle.Mutable.compile_string(s), f, 1)
        p1.tuples, p1.relax_calls = ls, la
        return p1
    return oracle.csp_bytes(f, s)


def verify_batch(torch, oracle, f, args, data, offsets, st, poff, d_il, d_ol, d_w, d_fin):
    """Every string of the timed batch: equal strings must have equal results (device-side comparison over the whole
    batch), and one string per class — 64 strings when all strings differ — must equal the oracle's path bit for bit.
    Returns (strings covered, oracle runs, mean oracle tuples, mean oracle relax calls, seconds per oracle run)."""
    eager = args.semantics == "eager"
    n = len(st)
    lens = np.diff(offsets.astype(np.int64))
    dev = d_il.device
    t_poff = torch.from_numpy(poff.astype(np.int64)).to(dev)
    covered, runs, tr, rr, t_or = 0, 0, 0, 0, 0.0

    def check_one(i):
        nonlocal runs, tr, rr, t_or
        a, b = int(offsets[i]), int(offsets[i + 1])
        t0 = time.time()
        p1 = oracle_path(oracle, f, data[a:b].tobytes(), eager)
        t_or += time.time() - t0
        lo, hi = int(poff[i]), int(poff[i + 1])
        ok = (st[i] == 0) == (p1.status == oracle.STATUS_OK)
        if ok and st[i] == 0:
            il = d_il[lo:hi].cpu().numpy().astype(np.uint32); ol = d_ol[lo:hi].cpu().numpy().astype(np.uint32); w = d_w[lo:hi].cpu().numpy()
            ok = (np.array_equal(il, p1.ilabels) and np.array_equal(ol, p1.olabels) and np.array_equal(w.view(np.uint64), p1.weights.view(np.uint64)) and
                  np.float64(d_fin[i].item()).view(np.uint64) == np.float64(p1.final_weight).view(np.uint64))
        assert ok, f"bench: GPU path of string {i} differs from the oracle"
        runs += 1; tr += p1.tuples; rr += p1.relax_calls

    if args.workload == "wetext":
        k = min(n, 64)
        for i in range(k):
            check_one(i)
        covered = k
    else:
        # strings of one length are identical: class representative vs oracle, every member vs the representative
        for length in np.unique(lens):
            members = np.flatnonzero(lens == length)
            rep = int(members[0])
            check_one(rep)
            P = int(poff[rep + 1] - poff[rep])
            m = torch.from_numpy(members).to(dev)
            assert bool(((t_poff[m + 1] - t_poff[m]) == P).all()), f"bench: path lengths differ inside the class of length {length}"
            assert bool((torch.from_numpy(st[members].astype(np.int64)) == int(st[rep])).all()), "bench: statuses differ inside a class"
            if P and st[rep] == 0:
                base = int(poff[rep])
                blk = max(1, (1 << 22) // P)
                for b0 in range(0, len(members), blk):
                    idx = t_poff[m[b0:b0 + blk]][:, None] + torch.arange(P, device=dev)[None, :]
                    for arr in (d_il, d_ol, d_w.view(torch.int64)):
                        assert bool((arr[idx] == arr[base:base + P][None, :]).all()), f"bench: results differ inside the class of length {length}"
            assert bool((d_fin[m].view(torch.int64) == d_fin[rep].view(torch.int64)).all())
            covered += len(members)
    return covered, runs, tr / max(runs, 1), rr / max(runs, 1), t_or / max(runs, 1)


def single_call_latency(L, oracle, fst, f, args, iters=20, warm=3):
    """BASELINE config 1 is the reference's per-call latency bench (bench/optimize-bench.zig:416-453: avg_ns over iters
    after warm-up): the same for the drop-in fst_compose_frozen_shortest_path on the GPU (compile_string outside the
    timed region, like :388-402) and for the CPU port on one core."""
    s = input_string(args.workload, args.len, args.branches)
    a = L.MutableFst.compile_string(s)
    for _ in range(warm):
        r = L.compose_frozen_shortest_path(a, fst, 1)
    t0 = time.perf_counter()
    for _ in range(iters):
        r = L.compose_frozen_shortest_path(a, fst, 1)
    gpu_ns = (time.perf_counter() - t0) / iters * 1e9
    assert r is not None
    lhs = oracle.Mutable.compile_string(s)
    for _ in range(warm):
        oracle.csp_mutable(lhs, f, 1)
    t0 = time.perf_counter()
    for _ in range(iters):
        p = oracle.csp_mutable(lhs, f, 1)
    cpu_ns = (time.perf_counter() - t0) / iters * 1e9
    il, ol, w, fw = r.chain()
    assert np.array_equal(il, p.ilabels) and np.array_equal(ol, p.olabels)
    return {"gpu_avg_ns": gpu_ns, "cpu_port_avg_ns": cpu_ns, "iters": iters, "warmup": warm,
            "note": "one string per call through fst_compose_frozen_shortest_path (a compiled string runs as a one-string batch: "
                    "one 8-lane group of the lean / fast kernel); CPU = C++ port of the reference, one core; the batch entry is the throughput path"}


def run_matrix(args):
    lines = []
    for c in MATRIX:
        cmd = [sys.executable, os.path.abspath(__file__), "--config", c, "--steps", str(args.steps), "--warmup", str(args.warmup)]
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
        out = [l for l in r.stdout.splitlines() if l.startswith("{")]
        line = out[-1] if out else json.dumps({"failed": c, "rc": r.returncode, "stderr": r.stderr[-600:]})
        sys.stderr.write(f"[matrix] config {c}: {time.time() - t0:.0f} s rc {r.returncode}\n")
        lines.append(line)
        emit(json.loads(line))
        if args.out:
            with open(args.out, "a") as fh:
                fh.write(line + "\n")


def run_strong(args):
    """ONE batch split over N GPUs by fst_compose_frozen_shortest_path_batch_multi (host buffers in, pinned host results
    out; one process, one host thread per GPU, no collective)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import libfst_b200 as L
    L.load()
    n_dev = L.device_count()
    if n_dev < args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but {n_dev} visible")
    L.configure(lanes_per_string=args.lanes, exhaustive=args.exhaustive, engine=args.engine, tuples_hint=args.tuples_hint,
                semantics=L.EAGER if args.semantics == "eager" else L.LAZY)
    devices = list(range(args.gpus))
    rflags = L.RESULT_NO_PATHS if args.bytes_only else 0
    total = args.total or {"epsilon_dense": 1 << 18, "ambiguous": 1 << 21, "plain": 1 << 22, "wetext": 10000000}[args.workload]
    fst, data, offsets, max_len, oracle_loader, sources = make_workload(args, total, seed=1)
    nbytes = int(offsets[-1])
    # warm-up: images uploaded, arenas sized, pinned result buffers in the pool
    for _ in range(max(1, min(args.warmup, 2))):
        L.compose_frozen_shortest_path_batch_multi(fst, data, offsets, devices=devices, chunks_per_device=args.chunks_per_device, copy=False, flags=rflags)
    samplers = [ClockSampler(d) for d in devices[:1]]
    for s in samplers:
        s.start()
    wall, dev_ms, launches, relax, tuples, d2h = 0.0, 0.0, 0, 0, 0, 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        m = L.compose_frozen_shortest_path_batch_multi(fst, data, offsets, devices=devices, chunks_per_device=args.chunks_per_device, copy=False, flags=rflags)
        wall += time.perf_counter() - t0
        dev_ms += m.device_ms; launches += m.launches; relax = m.total_relax; tuples = m.total_tuples; d2h = m.d2h_bytes
    for s in samplers:
        s.stop_flag.set(); s.join(timeout=2)
    # verification: a sample of the batch through the same entry with copies, against the oracle and the 1-GPU entry
    import oracle
    f = oracle_loader()
    k = min(total, 4096)
    sub_d, sub_o = data[:int(offsets[k])], offsets[:k + 1]
    msub = L.compose_frozen_shortest_path_batch_multi(fst, sub_d, sub_o, devices=devices, chunks_per_device=2).flat()
    one = L.compose_frozen_shortest_path_batch(fst, sub_d, sub_o)
    assert np.array_equal(msub.status, one.status) and np.array_equal(msub.ilabels, one.ilabels) and np.array_equal(msub.olabels, one.olabels)
    assert np.array_equal(msub.weights.view(np.uint64), one.weights.view(np.uint64)) and np.array_equal(msub.out_bytes, one.out_bytes)
    n_or = 0
    for i in range(0, k, max(1, k // 32)):
        p = oracle_path(oracle, f, sub_d[int(sub_o[i]):int(sub_o[i + 1])].tobytes(), args.semantics == "eager")
        if p.status == oracle.STATUS_OK:
            il, ol, w = msub.path(i)
            assert np.array_equal(il, p.ilabels) and np.array_equal(ol, p.olabels) and np.array_equal(w.view(np.uint64), p.weights.view(np.uint64))
        n_or += 1
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    relax_ps, tuples_ps = relax / total, tuples / total
    path_arcs = float(one.path_offsets[-1]) / k
    alg_bytes = 20.0 * relax_ps + 16.0 * tuples_ps + 4.0 * (nbytes / total) + 16.0 * path_arcs
    value = total * args.steps / (dev_ms / 1e3)
    e2e_v = total * args.steps / wall
    achieved = alg_bytes * total / (dev_ms / args.steps / 1e3) / 1e9 / args.gpus
    line = {
        "metric": "strings/sec batched compose_shortest_path", "value": value, "unit": "strings/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, total // args.gpus, 1 << 40), total_strings=total, chunks=len(m.chunk_first) - 1,
                       entry="fst_compose_frozen_shortest_path_batch_multi (one process, one host thread per GPU, no collective)",
                       result=("output strings + statuses + path lengths (FST_B200_RESULT_NO_PATHS)" if args.bytes_only else "full paths + output strings"),
                       distinct_strings=min(total, args.distinct) if args.workload == "wetext" else len(np.unique(np.diff(offsets.astype(np.int64))))),
        "composed_arcs_per_sec": value * relax_ps,
        "work_per_string": {"path_arcs": path_arcs, "tuples_run": tuples_ps, "relax_run": relax_ps, "mean_len": nbytes / total,
                            "checked_vs_oracle": n_or, "checked_vs_single_gpu_entry": k},
        "transducer": transducer_size(fst),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "per": "GPU (largest per-device device time)", "alg_bytes_per_string": alg_bytes},
        "e2e": {"value": e2e_v, "unit": "strings/s", "h2d_bytes_per_step": int(data.nbytes + offsets.nbytes), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": wall / args.steps * 1e3},
        "gpu_launches": launches, "clocks": samplers[0].summary() if samplers else None,
    }
    emit(line)


def main():
    global _JSON_FD
    args = parse()
    # stdout carries exactly one JSON line: anything native libraries print there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, ...) is sent to stderr instead
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.matrix:
        return run_matrix(args)
    if args.impl == "reference":
        return run_reference(args)
    if args.scaling == "strong":
        return run_strong(args)

    import torch
    import libfst_b200 as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the b200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L.load()
    L.configure(lanes_per_string=args.lanes, exhaustive=args.exhaustive, engine=args.engine, tuples_hint=args.tuples_hint,
                semantics=L.EAGER if args.semantics == "eager" else L.LAZY)

    batch = args.batch or DEFAULT_BATCH[args.workload]
    # every rank searches its own `batch` strings (weak scaling); the transducer is replicated per GPU
    probe = args.batch == 0 and args.workload in ("epsilon_dense", "ambiguous")
    fst, data, offsets, max_len, oracle_loader, sources = make_workload(args, 296 if probe else batch, seed=rank + 1)
    if probe:
        # identical strings finish together: a step is a whole number of full waves of the strings the device holds in
        # flight (learnt from the engine: the second call knows the search size and reports its resident capacity)
        for _ in range(2):
            r = L.compose_frozen_shortest_path_batch(fst, data, offsets)
        cap = max(296, L.last_occupancy()["capacity"])
        waves = 4 if args.mixed else (1 if float(r.n_tuples.mean()) >= 100000 else 8)
        batch = min(cap * waves, 1 << 20)
        data, offsets, max_len = workload_strings(args, batch, rank + 1)
    nbytes = int(offsets[-1])

    # ── device-resident inputs/outputs (torch owns the memory; the library gets raw pointers) ──
    dev = torch.device("cuda", local)
    h_bytes = torch.from_numpy(data.copy() if nbytes else np.zeros(1, np.uint8)).pin_memory()
    h_off = torch.from_numpy(offsets.astype(np.int64)).pin_memory()
    d_bytes, d_off = h_bytes.to(dev), h_off.to(dev)
    cap = (16 if args.workload == "wetext" else 2) * nbytes + 32 * batch + 1024
    d_status = torch.empty(batch, dtype=torch.int32, device=dev)
    d_poff = torch.empty(batch + 1, dtype=torch.int64, device=dev)
    d_il = torch.empty(cap, dtype=torch.int32, device=dev)
    d_ol = torch.empty(cap, dtype=torch.int32, device=dev)
    d_w = torch.empty(cap, dtype=torch.float64, device=dev)
    d_fin = torch.empty(batch, dtype=torch.float64, device=dev)
    d_nt = torch.empty(batch, dtype=torch.int32, device=dev)
    out = L.DeviceOut(d_status.data_ptr(), d_poff.data_ptr(), d_il.data_ptr(), d_ol.data_ptr(), d_w.data_ptr(), d_fin.data_ptr(),
                      d_nt.data_ptr(), cap)
    stream = torch.cuda.current_stream()

    def step_device():
        rc = L.lib().fst_b200_batch_device(fst.h, d_bytes.data_ptr(), d_off.data_ptr(), batch, max_len, C.byref(out), stream.cuda_stream)
        if rc != 0:
            raise RuntimeError(f"fst_b200_batch_device failed: FstError {rc}")
        return L.last_counters()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    tuples_per_string = float(d_nt.double().mean().item())
    tuples_max = int(d_nt.max().item())
    occ = L.last_occupancy()
    state_bytes = tuples_per_string * 16 * min(batch, occ["resident"] or batch)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if state_bytes <= (1 << 30) else None
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, relax, kernel_ms = 0, 0, 0.0
    elapsed_ms = 0.0
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1)            # evict L2 between timed steps (outside the timed region)
            torch.cuda.synchronize()
        ev0.record(stream)
        c = step_device()
        ev1.record(stream)
        torch.cuda.synchronize()
        elapsed_ms += ev0.elapsed_time(ev1)
        launches += c["launches"]; relax += c["relaxations"]; kernel_ms += c["device_ms"]
    barrier()
    sampler.stop_flag.set(); sampler.join(timeout=2)
    # what was timed: statuses and paths of the whole batch (verified below)
    st = d_status.cpu().numpy(); poff = d_poff.cpu().numpy().astype(np.int64)
    assert ((st == 0) | (st == 1)).all(), "bench: a string ended with an error status"
    path_arcs = float(np.diff(poff).mean())

    # ── end to end through the host-buffer C ABI ──
    e2e = None
    if not args.no_e2e:
        hb, ho = (h_bytes.numpy() if nbytes else np.zeros(0, np.uint8)), offsets
        for _ in range(1):
            L.compose_frozen_shortest_path_batch(fst, hb, ho)
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        e2e_dev_ms = 0.0
        for _ in range(args.steps):
            # the C-ABI call with host buffers: H2D of the strings, search, D2H of the whole result into pinned host
            # memory, and the release of that result (copy=False: no numpy copies of the result inside the timed region)
            r = L.compose_frozen_shortest_path_batch(fst, hb, ho, copy=False)
            e2e_dev_ms += r.device_ms
            d2h = r.d2h_bytes
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        # the host-buffer result is the device-resident one (one more call, untimed, with copies)
        r = L.compose_frozen_shortest_path_batch(fst, hb, ho)
        hi = int(poff[-1])
        assert np.array_equal(r.status, st) and np.array_equal(r.path_offsets.astype(np.int64), poff)
        assert np.array_equal(r.ilabels, d_il[:hi].cpu().numpy().astype(np.uint32)) and np.array_equal(r.olabels, d_ol[:hi].cpu().numpy().astype(np.uint32))
        e2e = {"ms": e2e_ms, "h2d": int(hb.nbytes + ho.nbytes), "d2h": int(d2h), "dev_ms": e2e_dev_ms}

    # ── max over ranks ──
    if dist is not None:
        t = torch.tensor([elapsed_ms, kernel_ms, e2e["ms"] if e2e else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, kernel_ms = float(t[0]), float(t[1])
        if e2e:
            e2e["ms"] = float(t[2])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    total_strings = batch * world * args.steps
    value = total_strings / (elapsed_ms / 1e3)
    ms_per_step = elapsed_ms / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # algorithmic bytes per string: 20 B per relaxation (one SoA arc record), 16 B per tuple
    # (dist + back-pointer), 4 B per input label, 16 B per emitted path arc — from the counters of the run itself
    relax_per_string = relax / (batch * args.steps)
    alg_bytes = 20.0 * relax_per_string + 16.0 * tuples_per_string + 4.0 * (nbytes / batch) + 16.0 * path_arcs
    per_launch_bytes = alg_bytes * batch
    kernel_ms_per_launch = kernel_ms / args.steps
    achieved = per_launch_bytes / (kernel_ms_per_launch / 1e3) / 1e9
    tr = measured_traffic(args)
    line = {
        "metric": "strings/sec batched compose_shortest_path", "value": value, "unit": "strings/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, batch, state_bytes, occ["resident"]),
        "composed_arcs_per_sec": value * relax_per_string,
        "work_per_string": {"path_arcs": path_arcs, "tuples_run": tuples_per_string, "tuples_max": tuples_max,
                            "relax_run": relax_per_string, "mean_len": nbytes / batch},
        "transducer": transducer_size(fst),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (tr["dram_bytes_per_string"] * batch if tr else None),
                     "traffic_source": (f"profiles/traffic.json ({tr.get('source', 'ncu')}): dram bytes per string of a separate ncu capture of this "
                                        f"workload x batch; not measured in this run") if tr else None,
                     "l2_bytes_per_string": tr.get("l2_bytes_per_string") if tr else None,
                     "peak_source": peak_src,
                     "kernel": "search kernel + ordered emit of one fst_b200_batch_device call (device time of the call; the search kernel is > 99 %)",
                     "alg_bytes_per_string": alg_bytes, "kernel_ms_per_launch": kernel_ms_per_launch},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
    }
    if e2e:
        line["e2e"] = {"value": total_strings / (e2e["ms"] / 1e3), "unit": "strings/s", "h2d_bytes_per_step": e2e["h2d"],
                       "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": e2e["ms"] / args.steps,
                       "device_ms_per_step": e2e["dev_ms"] / args.steps}
    if not args.no_cpu_baseline:
        import oracle   # checker + CPU baseline leg only
        f = oracle_loader()
        cores = os.cpu_count() or 1
        covered, runs, t_ref, r_ref, one = verify_batch(torch, oracle, f, args, data, offsets, st, poff, d_il, d_ol, d_w, d_fin)
        line["work_per_string"].update({"tuples_ref": t_ref, "relax_ref": r_ref, "checked_vs_oracle": covered, "oracle_runs": runs})
        sample = cpu_sample_size(args, cores, max(one, 1e-7), batch)
        secs = oracle.csp_batch_bytes(f, data[:int(offsets[sample])], offsets[:sample + 1], n_threads=cores,
                                      eager=args.semantics == "eager")["seconds"]
        line["cpu_baseline"] = {"value": sample / secs, "unit": "strings/s", "cores": cores, "kind": "port",
                                "sample": f"first {sample} strings of the same batch on {cores} threads ({secs:.1f} s); C++ "
                                          f"restatement of the reference (zig toolchain absent)"}
        if args.latency:
            line["single_call"] = single_call_latency(L, oracle, fst, f, args)
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
