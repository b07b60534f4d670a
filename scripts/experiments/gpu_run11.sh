#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_literal.py -m gpu -q -x -k "eager or lattice or layouts or per_call" > gpurun_out/r2_pytest_gpu11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu11.log
OUT=gpurun_out/r2_bfspf.jsonl
run() { name=$1; shift; echo "## $name: $*" >> gpurun_out/r2_bfspf.err; ( "$@" >> $OUT 2>> gpurun_out/r2_bfspf.err ) || echo "{\"failed\": \"$name\"}" >> $OUT; sed -i "\$s/^{/{\"variant\": \"$name\", /" $OUT; }
rm -f $OUT gpurun_out/r2_bfspf.err
V=$PWD/libfst_b200/variants
run cfg5_bfspf timeout 600 python bench.py --config 5 --steps 3 --warmup 3
run cfg5_nobfspf env LIBFST_B200_SO=$V/nobfspf.so timeout 600 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline
run cfg4 timeout 600 python bench.py --config 4 --steps 3 --warmup 3
run plain timeout 600 python bench.py --config plain --steps 3 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
for l in open('gpurun_out/r2_bfspf.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:22s} batch {d['config']['batch_per_gpu_per_step']:7d} {d['value']:12.1f} str/s e2e {d['e2e']['value']:12.1f} frac {d['roofline']['frac']:.3f} checked {d['work_per_string'].get('checked_vs_oracle')}")
PY
