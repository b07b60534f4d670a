#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
OUT=gpurun_out/r2_lanes.jsonl
run() { name=$1; shift; echo "## $name: $*" >> gpurun_out/r2_lanes.err; ( "$@" >> $OUT 2>> gpurun_out/r2_lanes.err ) || echo "{\"failed\": \"$name\"}" >> $OUT; sed -i "\$s/^{/{\"variant\": \"$name\", /" $OUT; }
rm -f $OUT gpurun_out/r2_lanes.err
run cfg4_l16 timeout 400 $B --config 4 --lanes 16
run cfg4_l32 timeout 400 $B --config 4 --lanes 32
run plain_l32 timeout 400 $B --config plain --lanes 32
python - <<'PY'
import json
for l in open('gpurun_out/r2_lanes.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:22s} batch {d['config']['batch_per_gpu_per_step']:7d} resident {d['config']['resident_strings_per_gpu']} {d['value']:12.1f} str/s frac {d['roofline']['frac']:.3f}")
PY
