#!/bin/bash
# A/B of programmatic dependent launch (CLICA_PDL): parity tests with it on, then the bench with it on / off
TAG=${1:-pdl}
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_gpu_graphed.py tests/test_gpu_step.py tests/test_gpu_mlp.py tests/test_gpu_lpnce.py -q -m gpu -x -p no:cacheprovider > $O/pytest_${TAG}.log 2>&1
echo "tests rc=$?"; tail -4 $O/pytest_${TAG}.log | cut -c1-300
for V in 1 0; do
  CLICA_PDL=$V timeout -k 10 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>$O/bench_pdl${V}_${TAG}.err | tail -1 > $O/bench_pdl${V}_${TAG}.json
  python - <<PY
import json
try:
    d = json.load(open("$O/bench_pdl${V}_${TAG}.json")); k = d["kernels"]; c = d["c3_strong"]; kc = c["kernels"]
    print("PDL=$V  C2 %.4f ms e2e %.4f (tc %.3f simt %.3f loss %.3f+%.3f sum %.3f)   C3 %.4f ms (tc %.3f)" % (d["ms_per_step"], d["e2e"]["ms_per_step"], k["encoder_gemm"]["tc_ms"], k["encoder_gemm"]["simt_ms"], k["loss_fwd"]["ms_per_step"], k["loss_bwd"]["ms_per_step"], k["sum_ms"], c["ms_per_step"], kc["encoder_gemm"]["tc_ms"]))
except Exception as e:
    print("  no bench json:", e)
PY
  tail -2 $O/bench_pdl${V}_${TAG}.err | cut -c1-300
done
