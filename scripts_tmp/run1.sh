python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "" mb10 mb12; do for lanes in 32; do
  if [ -n "$v" ]; then export LIBFST_B200_SO=$PWD/libfst_b200/variants/$v.so; else unset LIBFST_B200_SO; fi
  python bench.py --steps 1 --warmup 3 --batch 9472 --no-cpu-baseline --no-e2e --lanes $lanes > gpurun_out/v_${v}_$lanes.log 2>&1
  echo "variant=${v:-mb8} lanes=$lanes $(python -c "import json,sys; d=json.loads(open('gpurun_out/v_${v}_$lanes.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])" 2>&1 | tail -1)"
done; done
