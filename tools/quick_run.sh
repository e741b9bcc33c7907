#!/bin/bash
# Short GPU pass: parity tests, bench (C2), ncu launch list.  Usage (GPU box): bash tools/quick_run.sh <tag>
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests -q -m gpu -p no:cacheprovider -x -s > $O/pytest_gpu_${TAG}.log 2>&1
echo "pytest rc=$?"; tail -25 $O/pytest_gpu_${TAG}.log | cut -c1-400
timeout 240 python bench.py --steps 50 --warmup 10 2>$O/bench_${TAG}.err | tail -1 > $O/bench_${TAG}.json
echo "bench rc=$?"; tail -5 $O/bench_${TAG}.err; python - <<PY
import json
try:
    d = json.load(open("$O/bench_${TAG}.json"))
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"], "mode", d.get("step_mode"), "launches/step", d["launches_per_step"])
    print(json.dumps(d["kernels"]))
except Exception as e:
    print("no bench json:", e)
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench_stdout_${TAG}.log 2>&1
echo "ncu list rc=$?"
