#!/bin/bash
# multi-GPU pass on one box (one bench per reserve value in $RESERVES, default "0"):  bash tools/r2_multi_lean.sh <N> <tag>
N=${1:-2}; TAG=${2:-r2m}; RESERVES=${RESERVES:-0}
O=gpurun_out; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 10 240 $RUN --master-port 29701 tools/sharded_check.py > $O/sharded_check_n${N}_${TAG}.log 2>&1
echo "sharded_check rc=$?"; grep "world=" $O/sharded_check_n${N}_${TAG}.log | cut -c1-250; tail -3 $O/sharded_check_n${N}_${TAG}.log | cut -c1-300
for RES in $RESERVES; do
  CLICA_ALLREDUCE_SM_RESERVE=$RES timeout -k 10 400 $RUN --master-port 2971$((RES % 10)) bench.py --gpus $N --steps 20 --warmup 5 2>$O/bench_n${N}_res${RES}_${TAG}.err | tail -1 > $O/bench_n${N}_res${RES}_${TAG}.json
  echo "bench N=$N reserve=$RES rc=$?"; tail -3 $O/bench_n${N}_res${RES}_${TAG}.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.load(open("$O/bench_n${N}_res${RES}_${TAG}.json"))
    k = d["kernels"]
    print("C2 weak N=$N res=$RES: ms/step %.4f value %.3e e2e %.4f | gemm %.3f loss %.3f+%.3f nccl %.3f adam %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], k["encoder_gemm"]["ms_per_step"], k["loss_fwd"]["ms_per_step"], k["loss_bwd"]["ms_per_step"], k["nccl_ms"], k["adam_ms"]))
    c = d.get("c3_strong")
    if c:
        k = c["kernels"]
        print("C3 strong N=$N res=$RES: ms/step %.4f value %.3e e2e %.4f | gemm %.3f (simt %.3f) loss %.3f+%.3f nccl %.3f adam %.3f misc %.3f other %.3f" % (c["ms_per_step"], c["value"], c["e2e"]["ms_per_step"], k["encoder_gemm"]["ms_per_step"], k["encoder_gemm"]["simt_ms"], k["loss_fwd"]["ms_per_step"], k["loss_bwd"]["ms_per_step"], k["nccl_ms"], k["adam_ms"], k["misc_ms"], k["torch_other_ms"]))
except Exception as e:
    print("no bench json:", e)
PY
done
if [ -n "$SWEEP" ]; then
  timeout -k 10 300 $RUN --master-port 29731 tools/loss_sweep.py --quick --out $O/loss_sweep_n${N}_${TAG}.json > $O/loss_sweep_n${N}_${TAG}.log 2>&1
  echo "loss sweep rc=$?"; tail -4 $O/loss_sweep_n${N}_${TAG}.log | cut -c1-300
fi
