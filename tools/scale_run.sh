#!/bin/bash
# 8-GPU box: weak-scaling C2 at N=8,4 and strong-scaling C3 at N=8,1 (results under gpurun_out/)
for N in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_r1_c2_weak_n$N.json
  python -c "import json; d=json.load(open('gpurun_out/bench_r1_c2_weak_n$N.json')); print('c2 weak N=$N', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 20 --warmup 5 --workload c3 --scaling strong 2>&1 | tail -1 > gpurun_out/bench_r1_c3_strong_n8.json
python -c "import json; d=json.load(open('gpurun_out/bench_r1_c3_strong_n8.json')); print('c3 strong N=8', d['ms_per_step'], d['value'], d['e2e']['value'], d['kernels'])"
timeout 200 python bench.py --steps 10 --warmup 3 --workload c3 --scaling strong --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r1_c3_strong_n1.json
python -c "import json; d=json.load(open('gpurun_out/bench_r1_c3_strong_n1.json')); print('c3 N=1', d['ms_per_step'], d['value'], d['e2e']['value'], d['kernels'])"
