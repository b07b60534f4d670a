#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_scale.sh 1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --batch 18160 --tuples-hint 800000"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:csp_batch_fast_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r2_fast96 $CMD > gpurun_out/r2_fast96_ncu.log 2>&1
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r2_cfg4_e2e.json 2> gpurun_out/r2_cfg4_e2e.err
timeout 600 python bench.py --config 1 --steps 3 --warmup 3 > gpurun_out/r2_cfg1_e2e.json 2> gpurun_out/r2_cfg1_e2e.err
