export LIBFST_B200_DEBUG=1
for b in 14208 16576 37888; do
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --batch $b > gpurun_out/b96_$b.log 2>gpurun_out/b96_$b.err; echo rc=$?; tail -1 gpurun_out/b96_$b.err; cut -c1-200 gpurun_out/b96_$b.log
done
