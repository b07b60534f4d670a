CMD="python bench.py --workload wetext --steps 1 --warmup 3 --batch 75776 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:csp_batch_lean -s 3 -c 1 -o gpurun_out/prof_wetext -f $CMD > gpurun_out/ncuwt.log 2>&1
echo rc=$?; cut -c1-200 gpurun_out/plainwt.log
