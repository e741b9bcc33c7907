#!/bin/bash
# same-box A/B of two builds of the library: gpurun_ab_old.so (a previous commit's build) against the in-tree one
TAG=${1:-libab}
O=gpurun_out; mkdir -p $O
run() {   # name, env...
  local name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>$O/bench_${name}_${TAG}.err | tail -1 > $O/bench_${name}_${TAG}.json
  python - <<PY
import json
try:
    d = json.load(open("$O/bench_${name}_${TAG}.json")); k = d["kernels"]; c = d["c3_strong"]; kc = c["kernels"]
    print("%-10s C2 %.4f ms e2e %.4f (tc %.3f simt %.3f loss %.3f+%.3f sum %.3f launches %.0f) loss %.6f  C3 %.4f ms (tc %.3f) loss %.6f" % ("$name", d["ms_per_step"], d["e2e"]["ms_per_step"], k["encoder_gemm"]["tc_ms"], k["encoder_gemm"]["simt_ms"], k["loss_fwd"]["ms_per_step"], k["loss_bwd"]["ms_per_step"], k["sum_ms"], d["launches_per_step"], d["loss"], c["ms_per_step"], kc["encoder_gemm"]["tc_ms"], c["loss"]))
except Exception as e:
    print("  $name: no bench json:", e)
PY
}
NEW=cl-ica_b200/lib/libclica_sm100.so
cp $NEW /tmp/new.so
run new_chain1 CLICA_TC_CHAIN=1
run new_chain0 CLICA_TC_CHAIN=0
if [ -f gpurun_ab_old.so ]; then cp gpurun_ab_old.so $NEW; run old X=1; cp /tmp/new.so $NEW; fi
run new_chain1b CLICA_TC_CHAIN=1
