#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
OUT=gpurun_out/r2_waves.jsonl
run() { name=$1; shift; echo "## $name: $*" >> gpurun_out/r2_waves.err; ( "$@" >> $OUT 2>> gpurun_out/r2_waves.err ) || echo "{\"failed\": \"$name\"}" >> $OUT; sed -i "\$s/^{/{\"variant\": \"$name\", /" $OUT; }
rm -f $OUT gpurun_out/r2_waves.err
run eps96_w2 timeout 400 $B --batch 36320
run eps96_w4 timeout 400 $B --batch 72640
run eps96_w2p5 timeout 400 $B --batch 45400
run eps251_w3 timeout 400 $B --len 251 --batch 20208
python - <<'PY'
import json
for l in open('gpurun_out/r2_waves.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:22s} batch {d['config']['batch_per_gpu_per_step']:7d} resident {d['config']['resident_strings_per_gpu']} {d['value']:12.1f} str/s frac {d['roofline']['frac']:.3f}")
PY
