#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
OUT=gpurun_out/r2_hmod.jsonl
run() { name=$1; shift; echo "## $name: $*" >> gpurun_out/r2_hmod.err; ( "$@" >> $OUT 2>> gpurun_out/r2_hmod.err ) || echo "{\"failed\": \"$name\"}" >> $OUT; sed -i "\$s/^{/{\"variant\": \"$name\", /" $OUT; }
rm -f $OUT gpurun_out/r2_hmod.err
for r in 1 2 3 0; do run eps96_h$r env LIBFST_B200_HMOD=$r timeout 300 $B; done
for r in 1 2 3 0; do run eps251_h$r env LIBFST_B200_HMOD=$r timeout 300 $B --len 251; done
for r in 1 3; do run eps33_h$r env LIBFST_B200_HMOD=$r timeout 300 $B --len 33; done
for r in 1 3; do run eps128_h$r env LIBFST_B200_HMOD=$r timeout 300 $B --len 128; done
python - <<'PY'
import json
for l in open('gpurun_out/r2_hmod.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:22s} batch {d['config']['batch_per_gpu_per_step']:7d} resident {d['config']['resident_strings_per_gpu']} {d['value']:12.1f} str/s frac {d['roofline']['frac']:.3f}")
PY
