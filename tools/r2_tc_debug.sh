#!/bin/bash
# timing experiments on the chained GEMM kernel: which part of the epilogue the C2 step is sensitive to (results are wrong)
TAG=${1:-dbg}
O=gpurun_out; mkdir -p $O
for V in ${VARIANTS:-0 1 2 4 7}; do
  CLICA_TC_DEBUG=$V timeout -k 10 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-c3 2>$O/bench_dbg${V}_${TAG}.err | tail -1 > $O/bench_dbg${V}_${TAG}.json
  python - <<PY
import json
try:
    d = json.load(open("$O/bench_dbg${V}_${TAG}.json")); k = d["kernels"]
    print("DEBUG=$V  C2 %.4f ms (tc %.3f simt %.3f loss %.3f+%.3f sum %.3f)" % (d["ms_per_step"], k["encoder_gemm"]["tc_ms"], k["encoder_gemm"]["simt_ms"], k["loss_fwd"]["ms_per_step"], k["loss_bwd"]["ms_per_step"], k["sum_ms"]))
except Exception as e:
    print("  no bench json:", e)
PY
done
