timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
LIBFST_B200_DEBUG=1 timeout 300 python bench.py --workload wetext --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/wt.log 2>gpurun_out/wt.err; echo rc=$?; grep "pass" gpurun_out/wt.err | tail -4; cut -c1-200 gpurun_out/wt.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b96.log 2>gpurun_out/b96.err; echo rc=$?; cut -c1-200 gpurun_out/b96.log
