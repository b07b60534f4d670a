"""Small batches through every kernel family (run under compute-sanitizer memcheck)."""
import sys, random
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import libfst_b200 as L
import oracle as O
from common import random_rhs, random_string, frozen_pair, gen_image, assert_batch_matches_oracle, assert_batch_matches_eager_oracle
from libfst_b200 import synth
L.load()
rng = random.Random(5)
spec = random_rhs(rng, max_states=8)
fprod, forc, _ = frozen_pair(L, O, spec)
strings = [random_string(rng, max_len=8) for _ in range(24)]
for engine, lanes in ((0, 0), (2, 8), (3, 8), (3, 4), (2, 16), (3, 32), (5, 0), (6, 0), (1, 0)):
    L.configure(engine=engine, lanes_per_string=lanes)
    assert_batch_matches_oracle(L, O, fprod, forc, strings)
L.configure(semantics=L.EAGER)
assert_batch_matches_eager_oracle(L, O, fprod, forc, strings)
L.configure()
img = gen_image(O, 1, 256, 12)
f, o = L.Fst.from_image(img), O.Frozen.from_bytes(img)
ss = [bytes(k) for k in (0, 3, 9)]
assert_batch_matches_oracle(L, O, f, o, ss)
data, off = L.pack_strings(ss)
lat = L.compose_frozen_lattice_batch(f, data, off)
for i, s in enumerate(ss):
    st, ab, fin, il, ol, w, nx = O.compose_bytes(o, s).dump()
    g = lat.lattice(i)
    assert np.array_equal(g[0], ab) and np.array_equal(g[5], nx)
L.compose_frozen_shortest_path_pipeline(fprod, fprod, *L.pack_strings(strings))
m, sources = synth.wetext_style(K=400)
fw = m.freeze()
ws = synth.wetext_strings(sources, 24, seed=3, lo=0, hi=40)
r = L.compose_frozen_shortest_path_batch(fw, *L.pack_strings(ws))
assert (r.status == L.PATH).all()
a = L.MutableFst.compile_string(bytes(5))
assert L.compose_frozen_shortest_path(a, f, 1) is not None
L.teardown()
print("sanity ok")
