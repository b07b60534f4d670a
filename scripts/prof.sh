# ncu captures summarised under profiles/ (run from the repo root on the GPU box; each capture only after the same
# command has exited 0 without ncu)
CMD="python bench.py --len 33 --steps 2 --warmup 3 --batch 18944 --no-cpu-baseline --no-e2e --tuples-hint 300000"
$CMD > gpurun_out/plain33.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo launches rc=$?
$CMD > gpurun_out/plain33b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:csp_batch_lean -s 3 -c 1 -o gpurun_out/prof_final33 -f $CMD > gpurun_out/ncu33.log 2>&1
echo full rc=$?
cut -c1-160 gpurun_out/plain33.log
