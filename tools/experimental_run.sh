#!/bin/bash
# Round-2 opener: validate the kernels written after round 1's GPU budget was spent, then measure them.
#   bash tools/experimental_run.sh <tag>        (GPU box; results under gpurun_out/)
TAG=${1:-x}
O=gpurun_out; mkdir -p $O
CLICA_EXPERIMENTAL=1 timeout -k 10 300 python -m pytest tests/test_gpu_experimental.py -q -p no:cacheprovider -s > $O/pytest_experimental_${TAG}.log 2>&1
echo "experimental tests rc=$?"; tail -15 $O/pytest_experimental_${TAG}.log | cut -c1-300
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  CLICA_FUSED_MIXING=$1 CLICA_LPNCE_FAST=$2 timeout -k 10 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>$O/bench_mix$1_fast$2_${TAG}.err | tail -1 > $O/bench_mix$1_fast$2_${TAG}.json
  python -c "
import json; d=json.load(open('$O/bench_mix$1_fast$2_${TAG}.json')); k=d['kernels']; print('mixing=$1 fast=$2: ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'loss_fwd', round(k['loss_fwd']['ms_per_step'],4), k['loss_fwd']['fp32_pipe_frac'])"
done
CLICA_LPNCE_FAST=1 timeout -k 10 240 python tools/loss_sweep.py --out $O/loss_sweep_fast_${TAG}.json > $O/loss_sweep_fast_${TAG}.log 2>&1; tail -3 $O/loss_sweep_fast_${TAG}.log | cut -c1-300
CLICA_PACK_FUSED=1 CLICA_FUSED_MIXING=1 CLICA_LPNCE_FAST=1 timeout -k 10 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>$O/bench_all_${TAG}.err | tail -1 > $O/bench_all_${TAG}.json
python -c "
import json; d=json.load(open('$O/bench_all_${TAG}.json')); print('all experimental switches: ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'launches/step', d['launches_per_step'])"
