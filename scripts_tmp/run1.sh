python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for lanes in 32 16; do
  python bench.py --steps 1 --warmup 3 --batch 9472 --no-cpu-baseline --no-e2e --lanes $lanes > gpurun_out/v__$lanes.log 2>&1
  echo "lanes=$lanes $(python -c "import json,sys; d=json.loads(open('gpurun_out/v__$lanes.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])" 2>&1 | tail -1)"
done
