#!/bin/bash
# N-GPU tuning of the overlapped gradient all-reduce at BASELINE config 3 (strong scaling), then the full default bench
#   bash tools/r2_n8_tune.sh <N> <tag>
N=${1:-8}; TAG=${2:-n8t}
O=gpurun_out; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 10 240 $RUN --master-port 29701 tools/sharded_check.py > $O/sharded_check_n${N}_${TAG}.log 2>&1
echo "sharded_check rc=$?"; tail -2 $O/sharded_check_n${N}_${TAG}.log | cut -c1-200
i=0
for V in "CLICA_NOP=1" "CLICA_GRAD_BUCKET_MB=1000" "CLICA_ALLREDUCE_SM_RESERVE=8 NCCL_MAX_CTAS=8" "CLICA_GRAD_BUCKET_MB=12"; do
  i=$((i+1))
  env $V timeout -k 10 300 $RUN --master-port 2972$i bench.py --gpus $N --steps 20 --warmup 5 --workload c3 --scaling strong --no-cpu-baseline 2>$O/bench_tune${i}_n${N}_${TAG}.err | tail -1 > $O/bench_tune${i}_n${N}_${TAG}.json
  python - <<PY
import json
try:
    d = json.load(open("$O/bench_tune${i}_n${N}_${TAG}.json")); k = d["kernels"]
    print("%-50s C3 strong N=$N: %.4f ms  value %.3e e2e %.4f | gemm %.3f loss %.3f+%.3f nccl %.3f adam %.3f misc %.3f other %.3f" % ("$V", d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], k["encoder_gemm"]["ms_per_step"], k["loss_fwd"]["ms_per_step"], k["loss_bwd"]["ms_per_step"], k["nccl_ms"], k["adam_ms"], k["misc_ms"], k["torch_other_ms"]))
except Exception as e:
    print("$V: no bench json:", e)
PY
done
timeout -k 10 400 $RUN --master-port 29731 bench.py --gpus $N --steps 20 --warmup 5 2>$O/bench_n${N}_${TAG}.err | tail -1 > $O/bench_n${N}_${TAG}.json
python - <<PY
import json
try:
    d = json.load(open("$O/bench_n${N}_${TAG}.json")); k = d["kernels"]; c = d["c3_strong"]; kc = c["kernels"]
    print("full: C2 weak N=$N %.4f ms value %.3e e2e %.4f | gemm %.3f loss %.3f+%.3f nccl %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], k["encoder_gemm"]["ms_per_step"], k["loss_fwd"]["ms_per_step"], k["loss_bwd"]["ms_per_step"], k["nccl_ms"]))
    print("full: C3 strong N=$N %.4f ms value %.3e e2e %.4f | gemm %.3f loss %.3f+%.3f nccl %.3f adam %.3f misc %.3f" % (c["ms_per_step"], c["value"], c["e2e"]["ms_per_step"], kc["encoder_gemm"]["ms_per_step"], kc["loss_fwd"]["ms_per_step"], kc["loss_bwd"]["ms_per_step"], kc["nccl_ms"], kc["adam_ms"], kc["misc_ms"]))
except Exception as e:
    print("full: no bench json:", e)
PY
