#!/usr/bin/env python3
"""Aggregate an ncu SASS-level source page by CUDA source line.
usage: by_line.py <report.ncu-rep> <lib.so> <kernel mangled-name substring> [top]
Joins `ncu --page source --csv` (per-SASS-instruction counters) with
`nvdisasm -g` line info of the cubin inside the .so (needs -lineinfo)."""
import csv, os, re, subprocess, sys, tempfile, collections

rep, so, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# instruction index -> (file, line) for the kernel
lines, cur, inside = [], None, False
for l in sass:
    if l.startswith("\t.section\t.text."):
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ci, cs = H.index("Instructions Executed"), H.index("# Samples")
body = [r for r in rows[hdr + 1:] if len(r) == len(H)]
agg = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for k, r in enumerate(body):
    key = lines[k] if k < len(lines) else None
    a = agg[key]
    a[0] += int(r[ci] or 0); a[1] += int(r[cs] or 0)
    tot_i += int(r[ci] or 0); tot_s += int(r[cs] or 0)
print(f"sass instructions {len(body)} (disasm {len(lines)}); executed {tot_i}; samples {tot_s}")
print(f"{'file:line':32s} {'inst%':>7s} {'samples%':>9s}")
for key, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{str(key[0]) + ':' + str(key[1]) if key else '?':32s} {100.0 * i / max(tot_i, 1):7.2f} {100.0 * s / max(tot_s, 1):9.2f}")

# optional: aggregate by named line ranges of one file:  REGIONS="file:name:lo-hi,name:lo-hi"
import os as _os
reg = _os.environ.get("REGIONS")
if reg:
    fname, rest = reg.split(":", 1)
    rs = []
    for item in rest.split(","):
        n, r = item.split(":"); lo, hi = r.split("-"); rs.append((n, int(lo), int(hi)))
    ragg = collections.defaultdict(lambda: [0, 0])
    for key, (i, s) in agg.items():
        name = "other:" + (key[0] if key else "?")
        if key and key[0] == fname:
            for n, lo, hi in rs:
                if lo <= key[1] <= hi:
                    name = n; break
            else:
                name = "other:" + fname
        ragg[name][0] += i; ragg[name][1] += s
    print("\nby region:")
    for n, (i, s) in sorted(ragg.items(), key=lambda kv: -kv[1][0]):
        print(f"{n:40s} inst% {100.0*i/max(tot_i,1):6.2f}  samples% {100.0*s/max(tot_s,1):6.2f}")
