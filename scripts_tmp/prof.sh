CMD="python bench.py --len 33 --steps 2 --warmup 3 --batch 18944 --no-cpu-baseline --no-e2e --tuples-hint 300000"
$CMD > gpurun_out/plain33.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_g8.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo rc=$?
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/plain96.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -k regex:csp_batch_lean -s 8 -c 1 --csv --log-file gpurun_out/dram96.csv $CMD2 > gpurun_out/ncu96.log 2>&1
echo rc=$?
cut -c1-200 gpurun_out/plain96.log; tail -5 gpurun_out/dram96.csv | cut -c1-400
