#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_literal.py -m gpu -q -x -k "multi or result_without or literal_lengths or per_call" > gpurun_out/r2_pytest_gpu10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu10.log
B="python bench.py --steps 2 --warmup 3 --no-e2e"
OUT=gpurun_out/r2_keyfwd.jsonl
run() { name=$1; shift; echo "## $name: $*" >> gpurun_out/r2_keyfwd.err; ( "$@" >> $OUT 2>> gpurun_out/r2_keyfwd.err ) || echo "{\"failed\": \"$name\"}" >> $OUT; sed -i "\$s/^{/{\"variant\": \"$name\", /" $OUT; }
rm -f $OUT gpurun_out/r2_keyfwd.err
V=$PWD/libfst_b200/variants
run eps96_base timeout 300 $B --no-cpu-baseline
run eps96_keyfwd env LIBFST_B200_SO=$V/keyfwd.so timeout 400 $B
run eps251_keyfwd env LIBFST_B200_SO=$V/keyfwd.so timeout 300 $B --len 251 --no-cpu-baseline
run amb96_keyfwd env LIBFST_B200_SO=$V/keyfwd.so timeout 300 $B --workload ambiguous --no-cpu-baseline
run mixed_keyfwd env LIBFST_B200_SO=$V/keyfwd.so timeout 400 $B --mixed
python - <<'PY'
import json
for l in open('gpurun_out/r2_keyfwd.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:22s} batch {d['config']['batch_per_gpu_per_step']:7d} resident {d['config']['resident_strings_per_gpu']} {d['value']:12.1f} str/s frac {d['roofline']['frac']:.3f} checked {d['work_per_string'].get('checked_vs_oracle')}")
PY
