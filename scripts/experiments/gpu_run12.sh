#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:csp_batch_lean_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r2_wetext $CMD > gpurun_out/r2_wetext_ncu.log 2>&1
