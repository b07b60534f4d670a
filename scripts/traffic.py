#!/usr/bin/env python3
"""Measured DRAM / L2 traffic of the search kernel per BASELINE config -> profiles/traffic.json (feeds roofline.traffic
in bench.py).  Run on the GPU box:   python scripts/traffic.py [config ...]
For every config: one `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum`
pass (single-pass metrics, no replay) over `bench.py --config C --steps 1 --no-cpu-baseline --no-e2e`; the longest launch
of the search kernel is the timed step's; bytes are divided by the batch of that step (printed by bench.py).
The raw per-launch CSVs are kept under gpurun_out/ (copy the ones to be judged into profiles/)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

configs = sys.argv[1:] or ["1", "2", "3:33", "3:251", "3a:251", "mixed", "4", "5"]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "profiles", "traffic.json")
try:
    table = json.load(open(path))
except Exception:
    table = {}
for c in configs:
    tag = c.replace(":", "_")
    log = os.path.join(ROOT, "gpurun_out", f"r2_traffic_{tag}.csv")
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum", "--clock-control", "none",
           "-k", "regex:csp_batch", "--csv", "--log-file", log,
           sys.executable, os.path.join(ROOT, "bench.py"), "--config", c, "--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-e2e"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not lines:
        print("failed", c, r.stderr[-400:]); continue
    line = json.loads(lines[-1])
    batch = line["config"]["batch_per_gpu_per_step"]
    rows = [x for x in csv.reader(open(log)) if len(x) > 5]
    hdr = next(x for x in rows if "Metric Name" in x)
    iid, ik, im, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    launches = {}
    for x in rows[rows.index(hdr) + 1:]:
        launches.setdefault(x[iid], {"kernel": x[ik]})[x[im]] = float(x[iv].replace(",", ""))
    best = max(launches.values(), key=lambda d: d.get("gpu__time_duration.sum", 0))
    unit = 1.0
    args = bench.parse(["--config", c])
    dram = best["dram__bytes_read.sum"] + best["dram__bytes_write.sum"]
    table[bench.traffic_key(args)] = {
        "dram_bytes_per_string": dram * unit / batch, "l2_bytes_per_string": best["lts__t_bytes.sum"] / batch,
        "dram_read": best["dram__bytes_read.sum"], "dram_write": best["dram__bytes_write.sum"], "lts_bytes": best["lts__t_bytes.sum"],
        "kernel_ns_under_ncu": best["gpu__time_duration.sum"], "batch": batch, "kernel": best["kernel"][:80],
        "source": f"ncu single-pass metrics, gpurun_out/r2_traffic_{tag}.csv, bench.py --config {c} --steps 1"}
    print(c, table[bench.traffic_key(args)])
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
