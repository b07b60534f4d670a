# config matrix (BASELINE.json configs 1, 3, 4, 5 + mixed batch); each line appended to gpurun_out/matrix.jsonl
rm -f gpurun_out/matrix.jsonl
run() { echo "## $*" >> gpurun_out/matrix.err; timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" >> gpurun_out/matrix.jsonl 2>> gpurun_out/matrix.err || echo "{\"failed\": \"$*\"}" >> gpurun_out/matrix.jsonl; }
run --workload ambiguous --len 96
for l in 11 33 64 128 192 251; do run --workload epsilon_dense --len $l; done
for l in 11 33 160 251; do run --workload ambiguous --len $l; done
run --workload epsilon_dense --mixed
run --workload ambiguous --mixed
run --workload plain --len 96
run --workload wetext
run --workload ambiguous --len 251 --semantics eager
python - <<'PY'
import json
for l in open('gpurun_out/matrix.jsonl'):
    d=json.loads(l)
    if 'failed' in d: print(d); continue
    print(f"{d['config']['workload'][:95]:95s} batch {d['config']['batch_per_gpu_per_step']:7d} {d['value']:12.1f} str/s  {d['composed_arcs_per_sec']/1e9:7.2f} Garcs/s frac {d['roofline']['frac']:.3f} N {d['work_per_string']['tuples_run']:.0f}")
PY
