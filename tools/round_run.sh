#!/bin/bash
# One-GPU evidence pass for a round: bench line, GPU parity tests, ncu launch list of the bench command, one
# `ncu --set full` capture of the hot kernels, the config-5 loss sweep and the C3 single-GPU bench.
# Usage (GPU box): bash tools/round_run.sh <tag>      (everything lands in gpurun_out/)
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_${TAG}.txt 2>&1
timeout 240 python bench.py --steps 50 --warmup 10 2>$O/bench_${TAG}.err | tail -1 > $O/bench_${TAG}.json
echo "bench rc=$?"; head -c 600 $O/bench_${TAG}.json; echo
timeout 420 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/pytest_gpu_${TAG}.log 2>&1
echo "pytest rc=$?"; tail -5 $O/pytest_gpu_${TAG}.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench_stdout_${TAG}.log 2>&1
echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc|lpnce|skinny|gemm_simt' -s 80 -c 40 -f -o $O/prof_${TAG}_step \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_stdout_${TAG}.log 2>&1
echo "ncu full rc=$?"
timeout 240 python tools/loss_sweep.py --out $O/loss_sweep_${TAG}.json > $O/loss_sweep_${TAG}.log 2>&1
echo "sweep rc=$?"; tail -3 $O/loss_sweep_${TAG}.log
timeout 150 python bench.py --steps 10 --warmup 3 --workload c3 --scaling strong --no-cpu-baseline 2>$O/bench_${TAG}_c3.err | tail -1 > $O/bench_${TAG}_c3.json
echo "c3 rc=$?"; head -c 400 $O/bench_${TAG}_c3.json; echo
ls -la $O
