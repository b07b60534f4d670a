#!/bin/bash
# round-2 GPU call 8: the whole GPU suite, the per-config matrix (roofline + cpu_baseline + e2e + verification), measured
# DRAM/L2 traffic per config, and the ncu captures of the headline kernel
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_final.log
rm -f gpurun_out/r2_matrix.jsonl
timeout 3000 python bench.py --matrix --steps 3 --warmup 3 --out gpurun_out/r2_matrix.jsonl > gpurun_out/r2_matrix.stdout 2> gpurun_out/r2_matrix.err
timeout 1500 python scripts/traffic.py > gpurun_out/r2_traffic.log 2>&1
cp profiles/traffic.json gpurun_out/r2_traffic.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/r2_headline_plain.json 2> gpurun_out/r2_headline_plain.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_epsdense96.csv $CMD > gpurun_out/r2_launches_epsdense96.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:csp_batch_fast_kernel --launch-skip 4 -c 1 -f -o gpurun_out/r2_fast96 $CMD > gpurun_out/r2_fast96_ncu.log 2>&1
