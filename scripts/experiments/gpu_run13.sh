#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu13.log
timeout 600 python __graft_entry__.py > gpurun_out/r2_smoke13.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke13.log
timeout 600 python bench.py --config 1 --steps 3 --warmup 3 > gpurun_out/r2_cfg1_latency.json 2> gpurun_out/r2_cfg1_latency.err
