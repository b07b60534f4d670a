#!/bin/bash
# round-2 GPU call 1: sanitizer on the smoke test, the whole GPU suite, and the kernel variants on the headline workload
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py > gpurun_out/r2_sanitizer_smoke.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/r2_sanitizer_smoke.log
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; echo "## $name: $*" >> gpurun_out/r2_variants.err; ( "$@" >> gpurun_out/r2_variants.jsonl 2>> gpurun_out/r2_variants.err ) || echo "{\"failed\": \"$name\"}" >> gpurun_out/r2_variants.jsonl; sed -i "\$s/^{/{\"variant\": \"$name\", /" gpurun_out/r2_variants.jsonl; }
rm -f gpurun_out/r2_variants.jsonl gpurun_out/r2_variants.err
run leaf_eps96 timeout 300 $B
run old_eps96 env LIBFST_B200_NO_FAST=1 timeout 300 $B
run inl8_eps96 env LIBFST_B200_SO=$PWD/libfst_b200/variants/inl8.so timeout 300 $B
run inl7_eps96 env LIBFST_B200_SO=$PWD/libfst_b200/variants/inl7.so timeout 300 $B
run leafpf_eps96 env LIBFST_B200_SO=$PWD/libfst_b200/variants/leafpf.so timeout 300 $B
run leaf_amb96 timeout 300 $B --workload ambiguous
run old_amb96 env LIBFST_B200_NO_FAST=1 timeout 300 $B --workload ambiguous
run leaf_eps251 timeout 300 $B --len 251
run old_eps251 env LIBFST_B200_NO_FAST=1 timeout 300 $B --len 251
python - <<'PY'
import json
for l in open('gpurun_out/r2_variants.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'failed' in d: print(d); continue
    print(f"{d['variant']:14s} batch {d['config']['batch_per_gpu_per_step']:7d} resident {d['config']['resident_strings_per_gpu']} {d['value']:12.1f} str/s frac {d['roofline']['frac']:.3f}")
PY
