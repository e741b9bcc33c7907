#!/bin/bash
# A/B of the activation layout / converter variants of the tcgen05 GEMM: parity tests + C2 / C3 step time per variant
TAG=${1:-ab}
O=gpurun_out; mkdir -p $O
for V in "0 1" "1 1"; do
  set -- $V
  export CLICA_TC_SINGLE_PLANE=$1 CLICA_TC_CONV_TRUNC=$2
  timeout -k 10 300 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_step.py -q -m gpu -p no:cacheprovider > $O/pytest_mlp_sp$1_tr$2_${TAG}.log 2>&1
  echo "single_plane=$1 trunc=$2: mlp tests rc=$?"; tail -3 $O/pytest_mlp_sp$1_tr$2_${TAG}.log | cut -c1-300
  timeout -k 10 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>$O/bench_sp$1_tr$2_${TAG}.err | tail -1 > $O/bench_sp$1_tr$2_${TAG}.json
  python - <<PY
import json
try:
    d = json.load(open("$O/bench_sp$1_tr$2_${TAG}.json")); k = d["kernels"]; c = d["c3_strong"]; kc = c["kernels"]
    print("  C2 %.4f ms (tc %.3f simt %.3f)   C3 %.4f ms (tc %.3f simt %.3f)" % (d["ms_per_step"], k["encoder_gemm"]["tc_ms"], k["encoder_gemm"]["simt_ms"], c["ms_per_step"], kc["encoder_gemm"]["tc_ms"], kc["encoder_gemm"]["simt_ms"]))
except Exception as e:
    print("  no bench json:", e)
PY
done
