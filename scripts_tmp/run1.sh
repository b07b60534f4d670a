timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { timeout 400 python bench.py --steps 1 --warmup 3 --no-e2e "$@" > gpurun_out/tmp.log 2>gpurun_out/tmp.err; echo "$* => $(python -c "import json,sys; d=json.loads(open('gpurun_out/tmp.log').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],1), d['work_per_string'], d.get('cpu_baseline',{}).get('value'))" 2>&1 | tail -1)"; tail -2 gpurun_out/tmp.err; }
run --workload wetext --dict 20000 --batch 65536
run --workload wetext --dict 110000 --batch 262144
