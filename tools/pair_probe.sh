#!/bin/bash
# GPU pass: parity tests with the default GEMM path, then the CTA-pair (tcgen05 cta_group::2) GEMM: parity tests under
# a short timeout, per-shape timing against the single-CTA kernel, bench with pairs if parity is green.
TAG=${1:-p}
O=gpurun_out
mkdir -p $O
timeout -k 10 300 python -m pytest tests -q -m gpu -p no:cacheprovider -x > $O/pytest_gpu_${TAG}.log 2>&1
echo "pytest(default) rc=$?"; tail -4 $O/pytest_gpu_${TAG}.log | cut -c1-300
CLICA_TC_PAIR=1 timeout -k 10 150 python -m pytest tests/test_gpu_linear.py tests/test_gpu_mlp.py -q -p no:cacheprovider -x -s > $O/pytest_pair_${TAG}.log 2>&1
PRC=$?
echo "pytest(pair) rc=$PRC"; tail -15 $O/pytest_pair_${TAG}.log | cut -c1-300
nvidia-smi --query-gpu=name,clocks.sm,memory.used --format=csv,noheader
CLICA_TC_PAIR=0 timeout -k 10 120 python tools/gemm_bench.py > $O/gemm_bench_single_${TAG}.log 2>&1; tail -6 $O/gemm_bench_single_${TAG}.log
if [ $PRC -eq 0 ]; then
  CLICA_TC_PAIR=1 timeout -k 10 120 python tools/gemm_bench.py > $O/gemm_bench_pair_${TAG}.log 2>&1; tail -6 $O/gemm_bench_pair_${TAG}.log
  CLICA_TC_PAIR=1 timeout -k 10 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>$O/bench_pair_${TAG}.err | tail -1 > $O/bench_pair_${TAG}.json
  python -c "
import json; d=json.load(open('$O/bench_pair_${TAG}.json')); print('PAIR ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], json.dumps(d['kernels']['encoder_gemm']))"
fi
CLICA_TC_PAIR=0 timeout -k 10 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>$O/bench_single_${TAG}.err | tail -1 > $O/bench_single_${TAG}.json
python -c "
import json; d=json.load(open('$O/bench_single_${TAG}.json')); print('SINGLE ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], json.dumps(d['kernels']))"
