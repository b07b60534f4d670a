#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu17.log
B="python bench.py --steps 3 --warmup 3 --no-e2e"
OUT=gpurun_out/r2_lockstep.jsonl
run() { name=$1; shift; echo "## $name: $*" >> gpurun_out/r2_lockstep.err; ( "$@" >> $OUT 2>> gpurun_out/r2_lockstep.err ) || echo "{\"failed\": \"$name\"}" >> $OUT; sed -i "\$s/^{/{\"variant\": \"$name\", /" $OUT; }
rm -f $OUT gpurun_out/r2_lockstep.err
run cfg4_free timeout 400 $B --config 4
run cfg4_lockstep env LIBFST_B200_LOCKSTEP_HASH=1 timeout 400 $B --config 4 --no-cpu-baseline
run cfg4_free_l4 timeout 400 $B --config 4 --lanes 4 --no-cpu-baseline
python - <<'PY'
import json
for l in open('gpurun_out/r2_lockstep.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:22s} batch {d['config']['batch_per_gpu_per_step']:7d} resident {d['config']['resident_strings_per_gpu']} {d['value']:12.1f} str/s frac {d['roofline']['frac']:.3f} checked {d['work_per_string'].get('checked_vs_oracle')}")
PY
