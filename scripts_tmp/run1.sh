timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/tmp.log 2>&1; echo "$* => $(python -c "import json,sys; d=json.loads(open('gpurun_out/tmp.log').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],1))" 2>&1 | tail -1)"; }
run --batch 9472
run --batch 9472 --lanes 32
run --workload ambiguous --batch 65536
run --workload ambiguous --batch 65536 --exhaustive 1
run --workload plain --batch 1048576
