timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 2 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; tail -2 gpurun_out/bench_e2e.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'])"
