timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
LIBFST_B200_DEBUG=1 timeout 300 python bench.py --workload ambiguous --len 251 --semantics eager --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --batch 18944 > gpurun_out/eager.log 2>gpurun_out/eager.err; echo rc=$?; grep "pass 0" gpurun_out/eager.err | tail -1; cut -c1-200 gpurun_out/eager.log
