#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
OUT=gpurun_out/r2_minblocks.jsonl
run() { name=$1; shift; echo "## $name: $*" >> gpurun_out/r2_minblocks.err; ( "$@" >> $OUT 2>> gpurun_out/r2_minblocks.err ) || echo "{\"failed\": \"$name\"}" >> $OUT; sed -i "\$s/^{/{\"variant\": \"$name\", /" $OUT; }
rm -f $OUT gpurun_out/r2_minblocks.err
V=$PWD/libfst_b200/variants
run cfg4_mb8 timeout 400 $B --config 4
run cfg4_mb10 env LIBFST_B200_SO=$V/mb10.so timeout 400 $B --config 4
run cfg4_mb12 env LIBFST_B200_SO=$V/mb12.so timeout 400 $B --config 4
run plain_mb12 env LIBFST_B200_SO=$V/mb12.so timeout 400 $B --config plain
python - <<'PY'
import json
for l in open('gpurun_out/r2_minblocks.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:22s} batch {d['config']['batch_per_gpu_per_step']:7d} resident {d['config']['resident_strings_per_gpu']} {d['value']:12.1f} str/s frac {d['roofline']['frac']:.3f}")
PY
