#!/bin/bash
# Evidence pass: FP32-pipe calibration, config-4 conv bench, loss sweep (config 5, 1 GPU), ncu --set full of the C2 / C3
# GEMM + loss kernels.   bash tools/r2_evidence.sh <tag>
TAG=${1:-ev}
O=gpurun_out; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp32_pipe_probe tools/fp32_pipe_probe.cu 2>/dev/null
nvidia-smi --query-gpu=clocks.sm --format=csv,noheader,nounits -lms 50 > $O/fp32_probe_clocks_${TAG}.txt &
SMI_PID=$!
./tools/fp32_pipe_probe > $O/fp32_pipe_probe_${TAG}.log 2>&1; cat $O/fp32_pipe_probe_${TAG}.log
kill $SMI_PID 2>/dev/null; wait $SMI_PID 2>/dev/null
timeout -k 10 300 python tools/config4_bench.py --out $O/config4_${TAG}.json > $O/config4_${TAG}.log 2>&1; tail -1 $O/config4_${TAG}.log | cut -c1-600
timeout -k 10 400 python tools/loss_sweep.py --out $O/loss_sweep_${TAG}.json > $O/loss_sweep_${TAG}.log 2>&1; tail -2 $O/loss_sweep_${TAG}.log | cut -c1-300
timeout -k 10 420 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc|lpnce|skinny|adam|split_planes|mixing' -s 60 -c 28 -f -o $O/prof_c2_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c3 --no-graph > $O/ncu_full_c2_stdout_${TAG}.log 2>&1
echo "ncu c2 full rc=$?"
ncu -i $O/prof_c2_${TAG}.ncu-rep --page raw --csv > $O/prof_c2_${TAG}.csv 2>/dev/null; python tools/ncu_full_summary.py $O/prof_c2_${TAG}.csv > $O/prof_c2_${TAG}.md 2>/dev/null; head -40 $O/prof_c2_${TAG}.md | cut -c1-220
timeout -k 10 420 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc|lpnce|skinny|adam|split_planes|mixing' -s 60 -c 28 -f -o $O/prof_c3_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload c3 --scaling strong --no-graph > $O/ncu_full_c3_stdout_${TAG}.log 2>&1
echo "ncu c3 full rc=$?"
ncu -i $O/prof_c3_${TAG}.ncu-rep --page raw --csv > $O/prof_c3_${TAG}.csv 2>/dev/null; python tools/ncu_full_summary.py $O/prof_c3_${TAG}.csv > $O/prof_c3_${TAG}.md 2>/dev/null; head -40 $O/prof_c3_${TAG}.md | cut -c1-220
rm -f $O/prof_c2_${TAG}.ncu-rep $O/prof_c3_${TAG}.ncu-rep      # raw reports are large; the csv + md stay
