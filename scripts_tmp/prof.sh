set -x
CMD="python bench.py --len 33 --steps 2 --warmup 3 --batch 18944 --no-cpu-baseline --no-e2e --tuples-hint 300000"
$CMD > gpurun_out/plain33.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:csp_batch_lean -s 3 -c 1 -o gpurun_out/prof_lean33f -f $CMD > gpurun_out/ncu33.log 2>&1
cat gpurun_out/plain33.log | cut -c1-300
