import sys, os, tempfile, numpy as np
sys.path.insert(0, os.getcwd())
import libfst_b200 as L, oracle as O
from libfst_b200 import synth
L.load()
m, sources = synth.wetext_style(K=3000)
f = m.freeze()
p = tempfile.mktemp(); f.save(p); img = open(p, 'rb').read(); os.unlink(p)
fo = O.Frozen.from_bytes(img)
strings = synth.wetext_strings(sources, 160, seed=3, lo=0, hi=80)
for engine, lanes in ((0, 0), (2, 8), (1, 0)):
    L.configure(engine=engine, lanes_per_string=lanes)
    data, offsets = L.pack_strings(strings)
    res = L.compose_frozen_shortest_path_batch(f, data, offsets)
    bad = 0
    for i, s in enumerate(strings):
        r = O.csp_bytes(fo, s)
        il, ol, w = res.path(i)
        exp = r.output_bytes()
        got = res.output(i)
        rec = bytes(int(x) - 1 for x in ol if x)
        if got != exp:
            bad += 1
            if bad <= 3:
                print(engine, lanes, i, len(il), "exp", exp[:80], "\n   got", got[:80], "\n   rec==exp", rec == exp, len(got), len(exp),
                      res.out_offsets[i], res.out_offsets[i + 1], res.path_offsets[i], res.path_offsets[i+1])
    print("engine", engine, "bad", bad, "total path", res.path_offsets[-1], "out total", res.out_offsets[-1])
