#!/bin/bash
# Strong scaling of ONE batch over N GPUs through fst_compose_frozen_shortest_path_batch_multi (usage: gpu_scale.sh N).
# Run with `gpurun --gpus N`; N = 1 also runs the full-size single-GPU legs of the comparison.
N=${1:-1}
mkdir -p gpurun_out
OUT=gpurun_out/r2_strong_n$N${2:+_$2}.jsonl
ERR=gpurun_out/r2_strong_n$N${2:+_$2}.err
rm -f $OUT $ERR
nvidia-smi -L > gpurun_out/r2_gpus_n$N.txt 2>&1
run() { echo "## $*" >> $ERR; ( "$@" >> $OUT 2>> $ERR ) || echo "{\"failed\": \"$*\"}" >> $OUT; }
ONLY=${2:-all}
if [ "$ONLY" = all ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2_pytest_multi_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_multi_n$N.log
# config 2 (eps-dense len 96): one batch of 262 144 strings
run timeout 900 python bench.py --scaling strong --gpus $N --config 2 --total 262144 --steps 1 --warmup 1
fi
# config 4 (WeText-style tagger, ~1 M arcs): the literal 10 M-string batch (1 M distinct strings x 10), results = output
# strings + statuses + path lengths (the per-arc arrays of 10 M paths are 100 GB of D2H); and 1 M strings with full paths
run timeout 1200 python bench.py --scaling strong --gpus $N --config 4 --total 10000000 --distinct 1000000 --bytes-only --steps 1 --warmup 1
run timeout 900 python bench.py --scaling strong --gpus $N --config 4 --total 1000000 --distinct 1000000 --steps 1 --warmup 1
if [ "$ONLY" = all ]; then
# config 5 (eager lattice + shortest path, len 251): 262 144 strings
run timeout 900 python bench.py --scaling strong --gpus $N --config 5 --total 262144 --steps 1 --warmup 1
fi
python - <<PY
import json
for l in open('$OUT'):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    if 'failed' in d: print(d); continue
    print(f"N={d['n_gpus']} {d['config']['workload'][:60]:60s} {d['config'].get('result','')[:12]:12s} total {d['config']['total_strings']:9d} value {d['value']:12.1f} e2e {d['e2e']['value']:12.1f} str/s")
PY
