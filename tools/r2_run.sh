#!/bin/bash
# Round-2 GPU pass: parity suite, bench (C2 headline + c3_strong), launch lists for C2 and C3.
#   bash tools/r2_run.sh <tag> [skip_tests]
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
if [ -z "$2" ]; then
  timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/pytest_gpu_${TAG}.log 2>&1
  echo "pytest rc=$?"; tail -12 $O/pytest_gpu_${TAG}.log | cut -c1-400
fi
timeout -k 10 400 python bench.py --steps 50 --warmup 10 2>$O/bench_${TAG}.err | tail -1 > $O/bench_${TAG}.json
echo "bench rc=$?"; tail -5 $O/bench_${TAG}.err
python - <<PY
import json
try:
    d = json.load(open("$O/bench_${TAG}.json"))
    print("C2 ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "eager", round(d["e2e"].get("eager_dropin_ms_per_step",0),4), "launches/step", d["launches_per_step"])
    print(json.dumps(d["kernels"]))
    print("roofline", json.dumps(d["roofline"]))
    print("cpu", json.dumps(d["cpu_baseline"]))
    c = d.get("c3_strong")
    if c: print("C3 ms/step", round(c["ms_per_step"],4), "e2e", round(c["e2e"]["ms_per_step"],4), json.dumps(c["kernels"]))
except Exception as e:
    print("no bench json:", e)
PY
timeout -k 10 300 python bench.py --impl reference --steps 5 --warmup 1 2>$O/bench_ref_${TAG}.err | tail -1 > $O/bench_ref_${TAG}.json
python -c "
import json; d=json.load(open('$O/bench_ref_${TAG}.json')); print('reference arm', d['value'], d['ms_per_step'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])"
timeout -k 10 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c2_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c3 --no-graph > $O/ncu_c2_stdout_${TAG}.log 2>&1
echo "ncu c2 list rc=$?"
timeout -k 10 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload c3 --scaling strong --no-graph > $O/ncu_c3_stdout_${TAG}.log 2>&1
echo "ncu c3 list rc=$?"
