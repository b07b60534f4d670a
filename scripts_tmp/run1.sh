export LIBFST_B200_DEBUG=1
timeout 120 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload ambiguous --len 251 --semantics eager --batch 8 > gpurun_out/tmp.log 2>gpurun_out/tmp.err; echo rc=$?; tail -8 gpurun_out/tmp.err; cat gpurun_out/tmp.log | cut -c1-300
unset LIBFST_B200_DEBUG
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b96.log 2>gpurun_out/b96.err; echo rc=$?; tail -3 gpurun_out/b96.err; cut -c1-600 gpurun_out/b96.log
