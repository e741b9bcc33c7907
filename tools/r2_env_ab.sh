#!/bin/bash
# bench under a list of environment settings:  bash tools/r2_env_ab.sh <tag> "A=1 B=2" "A=3" ...   ("-" = defaults)
TAG=$1; shift
O=gpurun_out; mkdir -p $O
i=0
for V in "$@"; do
  i=$((i+1))
  if [ "$V" = "-" ]; then V="CLICA_NOP=1"; fi
  env $V timeout -k 10 240 python bench.py --steps 50 --warmup 10 --no-cpu-baseline ${BENCH_ARGS:---no-c3} 2>$O/bench_env${i}_${TAG}.err | tail -1 > $O/bench_env${i}_${TAG}.json
  python - <<PY
import json
try:
    d = json.load(open("$O/bench_env${i}_${TAG}.json")); k = d["kernels"]
    line = "%-40s C2 %.4f ms e2e %.4f (tc %.3f simt %.3f loss %.3f+%.3f sum %.3f) loss %.6f" % ("$V", d["ms_per_step"], d["e2e"]["ms_per_step"], k["encoder_gemm"]["tc_ms"], k["encoder_gemm"]["simt_ms"], k["loss_fwd"]["ms_per_step"], k["loss_bwd"]["ms_per_step"], k["sum_ms"], d["loss"])
    c = d.get("c3_strong")
    if c: line += "  C3 %.4f ms (tc %.3f) loss %.6f" % (c["ms_per_step"], c["kernels"]["encoder_gemm"]["tc_ms"], c["loss"])
    print(line)
except Exception as e:
    print("  $V: no bench json:", e)
PY
done
