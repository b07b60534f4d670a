#!/usr/bin/env python3
"""bench.py — batched compose_shortest_path throughput on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: oracle port of the reference, all host cores

One "step" = one pass of the hot path over one batch of synthetic strings of the
headline workload (BASELINE.json configs[1]): scenario
compose_frozen_lazy_shortest_path_epsilon_dense, len 96, transducer-len 4096,
branches 12.  The literal 1 M-string batch is processed as consecutive steps of
`--batch` strings per GPU (throughput is per string; the batch actually run is in
the JSON).  Weak scaling: every rank searches its own `--batch` strings against
its own replica of the transducer; there is no collective on the data path.

`value` : strings/s with inputs and outputs resident in HBM (fst_b200_batch_device).
`e2e`   : strings/s through fst_compose_frozen_shortest_path_batch with HOST buffers
          (H2D of the strings and D2H of the paths inside the timed region).
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="strings in the CPU baseline sample (0 = auto)")
    return ap.parse_args()


MIXED_LENS = [11, 19, 33, 64, 96, 128, 160, 192, 224, 251]   # bench/run_issue1_profile_bench.py:24-25
DEFAULT_BATCH = {"epsilon_dense": 9472, "ambiguous": 65536, "plain": 1 << 20, "wetext": 1 << 18}   # eps-dense: 64 strings per SM in flight


def input_string(workload: str, length: int, branches: int) -> bytes:
    if workload == "plain":
        return bytes(i % max(1, branches) for i in range(length))
    return bytes(length)


def workload_strings(args, batch, seed, sources=None):
    """(uint8 data, uint64 offsets, max_len) of `batch` input strings of the workload (numpy only)."""
    from libfst_b200 import synth
    if args.workload == "wetext":
        strings = synth.wetext_strings(sources, batch, seed=seed)
        lens = np.fromiter((len(x) for x in strings), np.uint64, len(strings))
        offsets = np.zeros(batch + 1, np.uint64); np.cumsum(lens, out=offsets[1:])
        data = np.frombuffer(b"".join(strings), np.uint8).copy()
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
    """Product arm: (Fst built through the C ABI, data, offsets, max_len, oracle_loader).  `oracle_loader()`
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
    return fst, data, offsets, max_len, oracle_loader


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
    sample = args.cpu_sample or int(max(cores, min(cores * 4096, cores * max(1.0, 5.0 / one))))
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
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
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
    return {"workload": desc, "semantics": args.semantics,
            "batch_per_gpu_per_step": batch, "resident_strings_per_gpu": resident, "literal_batch": 1000000,
            "cache": (f"per-step search state ~{state_bytes / 2**30:.1f} GiB in HBM >> 126 MB L2, rewritten by every string; "
                      f"no L2 flush needed") if state_bytes > (1 << 30) else
                     "search state fits L2: a buffer larger than L2 is written between timed steps",
            "lanes_per_string": args.lanes or "auto", "exhaustive": args.exhaustive, "engine": args.engine or "auto"}


def transducer_size(fst):
    """States and arcs of the frozen transducer, read back through the C ABI (fst_num_states / fst_num_arcs)."""
    n = int(fst.num_states())
    return {"states": n, "arcs": int(sum(fst.num_arcs(s) for s in range(n)))}


def measured_traffic(args):
    """DRAM bytes per string of the search kernel from the committed ncu capture of this workload (profiles/traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(f"{args.workload}:{args.len}:{args.transducer_len}:{args.branches}")
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


def main():
    global _JSON_FD
    args = parse()
    # stdout carries exactly one JSON line: anything native libraries print there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, ...) is sent to stderr instead
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

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
    fst, data, offsets, max_len, oracle_loader = make_workload(args, 296 if probe else batch, seed=rank + 1)
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
    # what was timed: statuses and paths of the first strings (checked against the oracle below)
    st = d_status.cpu().numpy(); poff = d_poff.cpu().numpy().astype(np.int64)
    assert ((st == 0) | (st == 1)).all(), "bench: a string ended with an error status"
    path_arcs = float(np.diff(poff).mean())
    n_check = min(batch, 32 if args.workload == "wetext" else 1)
    hi = int(poff[n_check])
    chk = (d_il[:hi].cpu().numpy().astype(np.uint32), d_ol[:hi].cpu().numpy().astype(np.uint32), d_w[:hi].cpu().numpy())

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
            r = L.compose_frozen_shortest_path_batch(fst, hb, ho)
            e2e_dev_ms += r.device_ms
            d2h = (r.status.nbytes + r.path_offsets.nbytes + r.ilabels.nbytes + r.olabels.nbytes + r.weights.nbytes +
                   r.final_weights.nbytes + r.n_tuples.nbytes + r.out_offsets.nbytes + r.out_bytes.nbytes)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
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
    traffic_ps = measured_traffic(args)
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
                     "traffic": (traffic_ps * batch if traffic_ps else None), "peak_source": peak_src,
                     "kernel": "csp_batch_lean_kernel" if not args.engine or args.engine >= 2 else "csp_batch_warp_kernel",
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
        # the timed GPU result must be the oracle's path, bit for bit (first strings of the batch)
        t0 = time.time()
        tr, rr = 0, 0
        for i in range(n_check):
            a, b = int(offsets[i]), int(offsets[i + 1])
            if args.semantics == "eager":
                p1, ls, la = oracle.eager_mutable(oracle.Mutable.compile_string(data[a:b].tobytes()), f, 1)
                p1.tuples, p1.relax_calls = ls, la
            else:
                p1 = oracle.csp_bytes(f, data[a:b].tobytes())
            lo, hi = int(poff[i]), int(poff[i + 1])
            ok = (st[i] == 0) == (p1.status == oracle.STATUS_OK)
            if ok and st[i] == 0:
                ok = (np.array_equal(chk[0][lo:hi], p1.ilabels) and np.array_equal(chk[1][lo:hi], p1.olabels) and
                      np.array_equal(chk[2][lo:hi].view(np.uint64), p1.weights.view(np.uint64)))
            assert ok, f"bench: GPU path of string {i} differs from the oracle"
            tr += p1.tuples; rr += p1.relax_calls
        one = max((time.time() - t0) / n_check, 1e-7)
        line["work_per_string"].update({"tuples_ref": tr / n_check, "relax_ref": rr / n_check, "checked_vs_oracle": n_check})
        sample = args.cpu_sample or int(max(cores, min(batch, cores * 4096, cores * max(1.0, 10.0 / one))))
        secs = oracle.csp_batch_bytes(f, data[:int(offsets[sample])], offsets[:sample + 1], n_threads=cores,
                                      eager=args.semantics == "eager")["seconds"]
        line["cpu_baseline"] = {"value": sample / secs, "unit": "strings/s", "cores": cores, "kind": "port",
                                "sample": f"first {sample} strings of the same batch on {cores} threads ({secs:.1f} s); C++ "
                                          f"restatement of the reference (zig toolchain absent)"}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
