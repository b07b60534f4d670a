#!/bin/bash
# round-2 GPU call 5: row-stream / child prefetch variants and the filter-halves row layout on the diagonal layout
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_literal.py tests/test_gpu_parity.py -m gpu -q -x -k "layouts or literal_lengths or engines_agree or eager or compact" > gpurun_out/r2_pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu5.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
OUT=gpurun_out/r2_pf.jsonl
run() { name=$1; shift; echo "## $name: $*" >> gpurun_out/r2_pf.err; ( "$@" >> $OUT 2>> gpurun_out/r2_pf.err ) || echo "{\"failed\": \"$name\"}" >> $OUT; sed -i "\$s/^{/{\"variant\": \"$name\", /" $OUT; }
rm -f $OUT gpurun_out/r2_pf.err
V=$PWD/libfst_b200/variants
run eps96_base timeout 300 $B
run eps96_noshift env LIBFST_B200_NOSHIFT=1 timeout 300 $B
run eps96_nochildpf env LIBFST_B200_SO=$V/nochildpf.so timeout 300 $B
run eps251 timeout 300 $B --len 251
run eps33 timeout 300 $B --len 33
run amb96_base timeout 300 $B --workload ambiguous
run eps128 timeout 300 $B --len 128
run eps192 timeout 300 $B --len 192
run mixed timeout 400 $B --mixed
python - <<'PY'
import json
for l in open('gpurun_out/r2_pf.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:22s} batch {d['config']['batch_per_gpu_per_step']:7d} resident {d['config']['resident_strings_per_gpu']} {d['value']:12.1f} str/s frac {d['roofline']['frac']:.3f}")
PY
