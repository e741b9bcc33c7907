#!/bin/bash
# loss-kernel pass: tests, timing probe (variants x regimes), one ncu --set full capture of the p = 2 kernels
TAG=${1:-r2d}
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_gpu_lpnce.py tests/test_gpu_kernels_r2.py -q -m gpu -p no:cacheprovider > $O/pytest_loss_${TAG}.log 2>&1
echo "loss tests rc=$?"; tail -8 $O/pytest_loss_${TAG}.log | cut -c1-400
timeout -k 10 300 python tools/loss_probe.py --out $O/loss_probe_c2_${TAG}.json > $O/loss_probe_c2_${TAG}.log 2>&1; cat $O/loss_probe_c2_${TAG}.log | cut -c1-250
timeout -k 10 300 python tools/loss_probe.py --B 768 --M 6144 --out $O/loss_probe_shard_${TAG}.json > $O/loss_probe_shard_${TAG}.log 2>&1; grep '"dot": "1", "r4": "1"' $O/loss_probe_shard_${TAG}.log | cut -c1-250
timeout -k 10 300 python tools/loss_probe.py --B 8192 --d 40 --p 3 --iters 10 > $O/loss_probe_c3_${TAG}.log 2>&1; cat $O/loss_probe_c3_${TAG}.log | cut -c1-250
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k 'regex:lpnce' -c 8 -f -o $O/prof_loss_${TAG} \
    python tools/loss_probe.py --once > $O/ncu_loss_stdout_${TAG}.log 2>&1
echo "ncu rc=$?"
ncu -i $O/prof_loss_${TAG}.ncu-rep --page raw --csv > $O/prof_loss_${TAG}.csv 2>/dev/null
python tools/ncu_full_summary.py $O/prof_loss_${TAG}.csv 2>/dev/null | head -20
