timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
LIBFST_B200_DEBUG=1 timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo rc=$?; grep -v "^\[libfst" gpurun_out/bench_default.err | tail -3; grep "pass 0" gpurun_out/bench_default.err | tail -1; cat gpurun_out/bench_default.json
