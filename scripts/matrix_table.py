#!/usr/bin/env python3
"""Markdown table of profiles/r2_matrix.jsonl."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [json.loads(l) for l in open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_matrix.jsonl"))]
tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
def fmt(v):
    return f"{v/1e6:.2f} M" if v >= 1e6 else (f"{v:,.0f}".replace(",", " ") if v >= 100 else f"{v:.1f}")
print("| config | workload (transducer-len 4096, branches 12) | batch (strings resident) | strings/s, inputs in HBM | strings/s end to end | CPU port strings/s | G composed arcs/s | `roofline.frac` | strings verified | DRAM / L2 MB per string (ncu) |")
print("|---|---|---|---|---|---|---|---|---|---|")
for d in rows:
    if "failed" in d:
        print(f"| {d['failed']} | FAILED | | | | | | | | |"); continue
    c, w = d["config"], d["work_per_string"]
    name = c["workload"].replace("compose_frozen_lazy_shortest_path_", "lazy ").replace("compose_frozen_lazy_shortest_path", "lazy plain").replace("compose_frozen_", "eager ")
    name = name.split(" transducer_len")[0].split(" dict=")[0]
    import bench
    a = bench.parse(["--config", c["baseline_config"]]) if c["baseline_config"] else None
    t = tr.get(bench.traffic_key(a)) if a else None
    ts = f"{t['dram_bytes_per_string']/1e6:.1f} / {t['l2_bytes_per_string']/1e6:.1f}" if isinstance(t, dict) else "-"
    print(f"| {c['baseline_config']} | {name} | {c['batch_per_gpu_per_step']:,} ({c['resident_strings_per_gpu']:,}) | {fmt(d['value'])} | {fmt(d['e2e']['value'])} | {fmt(d['cpu_baseline']['value'])} | "
          f"{d['composed_arcs_per_sec']/1e9:.1f} | {d['roofline']['frac']:.3f} | {w.get('checked_vs_oracle', 0):,} | {ts} |".replace(",", " "))
for d in rows:
    if d.get("single_call"):
        s = d["single_call"]; print(f"\nsingle call (config {d['config']['baseline_config']}): GPU {s['gpu_avg_ns']/1e6:.1f} ms, CPU port one core {s['cpu_port_avg_ns']/1e6:.1f} ms")
