#!/bin/bash
# Evidence pass with the CTA-pair GEMM as default: parity, bench (pair policies 1 and 2), launch list, full capture.
TAG=${1:-r1d}
O=gpurun_out
mkdir -p $O
timeout -k 10 300 python -m pytest tests -q -m gpu -p no:cacheprovider -x > $O/pytest_gpu_${TAG}.log 2>&1
echo "pytest rc=$?"; tail -4 $O/pytest_gpu_${TAG}.log | cut -c1-300
timeout -k 10 240 python bench.py --steps 50 --warmup 10 2>$O/bench_${TAG}.err | tail -1 > $O/bench_${TAG}.json
python -c "
import json; d=json.load(open('$O/bench_${TAG}.json')); print('PAIR=1 ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], json.dumps(d['kernels']))"
CLICA_TC_PAIR=2 timeout -k 10 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>$O/bench_pair2_${TAG}.err | tail -1 > $O/bench_pair2_${TAG}.json
python -c "
import json; d=json.load(open('$O/bench_pair2_${TAG}.json')); print('PAIR=2 ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], json.dumps(d['kernels']['encoder_gemm']))"
timeout -k 10 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench_stdout_${TAG}.log 2>&1
echo "ncu list rc=$?"
timeout -k 10 420 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc|lpnce|skinny|gemm_simt|adam|split_planes|colsum' -s 120 -c 30 -f -o $O/prof_${TAG}_step \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_full_stdout_${TAG}.log 2>&1
echo "ncu full rc=$?"
timeout -k 10 150 python bench.py --steps 10 --warmup 3 --workload c3 --scaling strong --no-cpu-baseline 2>$O/bench_${TAG}_c3.err | tail -1 > $O/bench_${TAG}_c3.json
python -c "
import json; d=json.load(open('$O/bench_${TAG}_c3.json')); print('C3 ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], json.dumps(d['kernels']['encoder_gemm']))"
