#!/bin/bash
# round-2 GPU call 2: the whole GPU suite + one ncu --set full capture of the fast kernel (eps-dense len 33, one full wave)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
CMD="python bench.py --len 33 --steps 2 --warmup 3 --batch 18944 --tuples-hint 300000 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/r2_fast33_bench.json 2> gpurun_out/r2_fast33_bench.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:csp_batch_fast_kernel -c 1 -f -o gpurun_out/r2_fast33 $CMD > gpurun_out/r2_fast33_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_fast33.csv $CMD > gpurun_out/r2_launches_fast33.log 2>&1
