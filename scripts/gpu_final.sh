#!/bin/bash
# round-2 final measurement call: the whole GPU suite, smoke, the per-config matrix, measured traffic, the default bench
# line and the CPU reference arm, all with the committed build
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_final.log
timeout 600 python __graft_entry__.py > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
rm -f gpurun_out/r2_matrix.jsonl
timeout 3000 python bench.py --matrix --steps 3 --warmup 3 --out gpurun_out/r2_matrix.jsonl > gpurun_out/r2_matrix.stdout 2> gpurun_out/r2_matrix.err
timeout 1500 python scripts/traffic.py > gpurun_out/r2_traffic.log 2>&1
cp profiles/traffic.json gpurun_out/r2_traffic.json
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
