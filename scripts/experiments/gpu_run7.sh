#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_literal.py tests/test_gpu_parity.py -m gpu -q -x -k "layouts or literal_lengths or eager or lattice" > gpurun_out/r2_pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu7.log
B="python bench.py --steps 2 --warmup 3 --no-e2e"
OUT=gpurun_out/r2_hrow.jsonl
run() { name=$1; shift; echo "## $name: $*" >> gpurun_out/r2_hrow.err; ( "$@" >> $OUT 2>> gpurun_out/r2_hrow.err ) || echo "{\"failed\": \"$name\"}" >> $OUT; sed -i "\$s/^{/{\"variant\": \"$name\", /" $OUT; }
rm -f $OUT gpurun_out/r2_hrow.err
run eps96_h97 timeout 400 $B
run eps96_h98 env LIBFST_B200_HPAD=1 timeout 300 $B --no-cpu-baseline
run eps96_h98s env LIBFST_B200_SHIFT=1 timeout 300 $B --no-cpu-baseline
run eps96_h99 env LIBFST_B200_HPAD=2 timeout 300 $B --no-cpu-baseline
run cfg5_fast timeout 600 $B --config 5
run cfg5_old env LIBFST_B200_NO_FAST=1 timeout 600 $B --config 5 --no-cpu-baseline
run amb251_fast timeout 400 $B --config 3a:251 --no-cpu-baseline
python - <<'PY'
import json
for l in open('gpurun_out/r2_hrow.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:22s} batch {d['config']['batch_per_gpu_per_step']:7d} resident {d['config']['resident_strings_per_gpu']} {d['value']:12.1f} str/s frac {d['roofline']['frac']:.3f} checked {d['work_per_string'].get('checked_vs_oracle')}")
PY
