#!/bin/bash
# round-2 GPU call 3: diagonal dense-table layout — parity, then the layout on the headline workload and the length sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_literal.py -m gpu -q -x -k "layouts or literal_lengths" > gpurun_out/r2_pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu3.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
OUT=gpurun_out/r2_layout.jsonl
run() { name=$1; shift; echo "## $name: $*" >> gpurun_out/r2_layout.err; ( "$@" >> $OUT 2>> gpurun_out/r2_layout.err ) || echo "{\"failed\": \"$name\"}" >> $OUT; sed -i "\$s/^{/{\"variant\": \"$name\", /" $OUT; }
rm -f $OUT gpurun_out/r2_layout.err
run eps96_auto timeout 300 $B
run eps96_rows env LIBFST_B200_SKEW=-1 timeout 300 $B
run eps96_skew1 env LIBFST_B200_SKEW=1 timeout 300 $B
run eps96_rows_3q env LIBFST_B200_SKEW=-1 timeout 300 $B --batch 14208
run eps251_rows env LIBFST_B200_SKEW=-1 timeout 300 $B --len 251
run eps96_auto_old env LIBFST_B200_NO_FAST=1 timeout 300 $B
run eps33_auto timeout 300 $B --len 33
run eps251_auto timeout 300 $B --len 251
run eps128_auto timeout 300 $B --len 128
run mixed_auto timeout 400 $B --mixed
run mixed_rows env LIBFST_B200_SKEW=-1 timeout 400 $B --mixed
run eps96_auto_half timeout 300 $B --batch 9472
run eps96_auto_3q timeout 300 $B --batch 14208
python - <<'PY'
import json
for l in open('gpurun_out/r2_layout.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:16s} batch {d['config']['batch_per_gpu_per_step']:7d} resident {d['config']['resident_strings_per_gpu']} {d['value']:12.1f} str/s frac {d['roofline']['frac']:.3f}")
PY
