"""BASELINE config 2 at its literal size: ONE batched call over 1 000 000 strings (host buffers in, host results out)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import libfst_b200 as L
from libfst_b200 import synth
L.load()
fst = synth.TRANSDUCERS["epsilon_dense"](4096, 12).freeze()
s = bytes(96)
def batch(n):
    return np.frombuffer(s * n, np.uint8), np.arange(n + 1, dtype=np.uint64) * len(s)
for n in (296, 296):                       # learn the search size (adaptive passes) before the big call
    L.compose_frozen_shortest_path_batch(fst, *batch(n))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
data, off = batch(n)
t0 = time.perf_counter()
r = L.compose_frozen_shortest_path_batch(fst, data, off)
dt = time.perf_counter() - t0
ok = bool((r.status == L.PATH).all())
first = r.output(0)
same = all(r.output(i) == first for i in range(0, n, max(1, n // 1000)))
po = r.path_offsets.astype(np.int64)
print(json.dumps({"literal_batch": n, "seconds": dt, "strings_per_s": n / dt, "device_ms": r.device_ms, "all_path": ok,
                  "outputs_identical_sampled": same, "path_arcs_total": int(po[-1]), "passes": int(r.passes),
                  "launches": int(r.launches), "composed_arcs": int(r.total_relax), "occupancy": L.last_occupancy()}))
