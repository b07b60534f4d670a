#!/usr/bin/env python3
"""Regenerates tests/golden/*.json from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  bench_signatures.json rows were first
cross-checked against SURVEY.md Appendix B (independent Python restatement):
N, R, P, total, final tuple and sha16 all agree."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle as O  # noqa: E402
from common import random_rhs, random_string  # noqa: E402


def main():
    rows = []
    for kind, lens in ((O.KIND_AMBIGUOUS, [11, 19, 33, 64, 96, 251]), (O.KIND_PLAIN, [11, 96, 251]), (O.KIND_EPS_DENSE, [11, 19, 33, 96, 251])):
        f = O.Frozen.generate(kind, 4096, 12)
        for L in lens:
            s = bytes(i % 12 for i in range(L)) if kind == O.KIND_PLAIN else bytes(L)
            p = O.csp_bytes(f, s)
            rows.append(dict(kind=kind, T=4096, B=12, L=L, N=p.tuples, R=p.relax_calls, P=len(p.ilabels), total=p.total,
                             final_tuple=list(p.final_tuple), sha16=p.signature(), pushes=p.pushes, retakes=p.retakes))
    json.dump(dict(note="oracle outputs; rows with L in SURVEY App. B match it exactly", rows=rows),
              open(os.path.join(HERE, "bench_signatures.json"), "w"), indent=1)

    cases = []
    for seed in range(40):
        rng = random.Random(1000 + seed)
        spec = random_rhs(rng)
        f = spec.to_oracle(O).freeze()
        strings = [random_string(rng) for _ in range(12)]
        paths = []
        for s in strings:
            p = O.csp_bytes(f, s)
            paths.append(None if p.status != O.STATUS_OK else
                         [list(map(int, p.ilabels)), list(map(int, p.olabels)), [float(x) for x in p.weights], p.final_weight])
        cases.append(dict(seed=1000 + seed, num_states=spec.num_states, finals=spec.finals, arcs=[list(a) for a in spec.arcs],
                          strings=[s.hex() for s in strings], paths=paths))
    json.dump(dict(note="tie-heavy random cases (SURVEY App. C generator); paths from the oracle", cases=cases),
              open(os.path.join(HERE, "fuzz_paths.json"), "w"))


if __name__ == "__main__":
    main()
