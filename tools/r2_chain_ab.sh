#!/bin/bash
# A/B of the chained GEMM launch (CLICA_TC_CHAIN): encoder parity tests with it on, then the bench with it on / off
TAG=${1:-chain}
O=gpurun_out; mkdir -p $O
timeout -k 10 420 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_linear.py tests/test_gpu_step.py tests/test_gpu_graphed.py -q -m gpu -x -p no:cacheprovider > $O/pytest_${TAG}.log 2>&1
echo "tests rc=$?"; tail -15 $O/pytest_${TAG}.log | cut -c1-300
for V in ${VARIANTS:-1 0}; do
  CLICA_TC_CHAIN=$V timeout -k 10 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>$O/bench_chain${V}_${TAG}.err | tail -1 > $O/bench_chain${V}_${TAG}.json
  python - <<PY
import json
try:
    d = json.load(open("$O/bench_chain${V}_${TAG}.json")); k = d["kernels"]; c = d["c3_strong"]; kc = c["kernels"]
    print("CHAIN=$V  C2 %.4f ms e2e %.4f (tc %.3f simt %.3f loss %.3f+%.3f sum %.3f launches %.0f) loss %.6f  C3 %.4f ms (tc %.3f) loss %.6f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], k["encoder_gemm"]["tc_ms"], k["encoder_gemm"]["simt_ms"], k["loss_fwd"]["ms_per_step"], k["loss_bwd"]["ms_per_step"], k["sum_ms"], d["launches_per_step"], d["loss"], c["ms_per_step"], kc["encoder_gemm"]["tc_ms"], c["loss"]))
except Exception as e:
    print("  no bench json:", e)
PY
  tail -2 $O/bench_chain${V}_${TAG}.err | cut -c1-300
done
