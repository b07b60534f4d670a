#!/bin/bash
# round-2 GPU call 4: skew sweep of the diagonal layout + next-pop speculation variants
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
OUT=gpurun_out/r2_skew.jsonl
run() { name=$1; shift; echo "## $name: $*" >> gpurun_out/r2_skew.err; ( "$@" >> $OUT 2>> gpurun_out/r2_skew.err ) || echo "{\"failed\": \"$name\"}" >> $OUT; sed -i "\$s/^{/{\"variant\": \"$name\", /" $OUT; }
rm -f $OUT gpurun_out/r2_skew.err
SPEC=$PWD/libfst_b200/variants/spec.so
SPECL1=$PWD/libfst_b200/variants/specl1.so
run eps96_s1 env LIBFST_B200_SKEW=1 timeout 300 $B
run eps96_s2 env LIBFST_B200_SKEW=2 timeout 300 $B
run eps96_s1_spec env LIBFST_B200_SKEW=1 LIBFST_B200_SO=$SPEC timeout 300 $B
run eps96_s1_specl1 env LIBFST_B200_SKEW=1 LIBFST_B200_SO=$SPECL1 timeout 300 $B
run eps96_rows_spec env LIBFST_B200_SKEW=-1 LIBFST_B200_SO=$SPEC timeout 300 $B
run eps251_s1 env LIBFST_B200_SKEW=1 timeout 300 $B --len 251
run eps251_s1_spec env LIBFST_B200_SKEW=1 LIBFST_B200_SO=$SPEC timeout 300 $B --len 251
run eps33_s1 env LIBFST_B200_SKEW=1 timeout 300 $B --len 33
run mixed_s1 env LIBFST_B200_SKEW=1 timeout 400 $B --mixed
run amb96_rows env LIBFST_B200_SKEW=-1 timeout 300 $B --workload ambiguous
run amb96_s0 env LIBFST_B200_SKEW=0 timeout 300 $B --workload ambiguous
run amb96_s1 env LIBFST_B200_SKEW=1 timeout 300 $B --workload ambiguous
run amb96_rows_spec env LIBFST_B200_SKEW=-1 LIBFST_B200_SO=$SPEC timeout 300 $B --workload ambiguous
run amb251_rows env LIBFST_B200_SKEW=-1 timeout 300 $B --workload ambiguous --len 251
run amb251_s0 env LIBFST_B200_SKEW=0 timeout 300 $B --workload ambiguous --len 251
python - <<'PY'
import json
for l in open('gpurun_out/r2_skew.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:18s} batch {d['config']['batch_per_gpu_per_step']:7d} resident {d['config']['resident_strings_per_gpu']} {d['value']:12.1f} str/s frac {d['roofline']['frac']:.3f}")
PY
