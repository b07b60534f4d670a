#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
OUT=gpurun_out/r2_keypf.jsonl
run() { name=$1; shift; ( "$@" >> $OUT 2>> gpurun_out/r2_keypf.err ) || echo "{\"failed\": \"$name\"}" >> $OUT; sed -i "\$s/^{/{\"variant\": \"$name\", /" $OUT; }
rm -f $OUT gpurun_out/r2_keypf.err
V=$PWD/libfst_b200/variants
run eps96_nolinepf env LIBFST_B200_SO=$V/nolinepf.so timeout 200 $B
run eps96_rollpf env LIBFST_B200_SO=$V/rollpf.so timeout 200 $B
python - <<'PY'
import json
for l in open('gpurun_out/r2_keypf.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:22s} {d['value']:12.1f} str/s frac {d['roofline']['frac']:.3f}")
PY
