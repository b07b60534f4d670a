timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>> gpurun_out/m2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(f\"{d['config']['workload'][:95]:95s} batch {d['config']['batch_per_gpu_per_step']:7d} {d['value']:12.1f} str/s  {d['composed_arcs_per_sec']/1e9:7.2f} Garcs/s frac {d['roofline']['frac']:.3f}\")"; }
run --workload epsilon_dense --mixed
run --workload ambiguous --mixed
run --workload ambiguous --len 96
run --workload ambiguous --len 251
run --workload wetext
