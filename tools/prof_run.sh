#!/bin/bash
# One-GPU profiling pass: bench line, ncu launch list of the same command, full captures of the top kernels.
# Usage (GPU box): bash tools/prof_run.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 30 --warmup 5 2>gpurun_out/bench_${TAG}.err | tail -1 > gpurun_out/bench_${TAG}.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_stdout_${TAG}.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 20 -c 6 -f -o gpurun_out/prof_${TAG}_gemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lpnce_ -c 4 -f -o gpurun_out/prof_${TAG}_loss \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
