#!/bin/bash
# N-GPU pass (default 2): sharded-step equivalence (eager and CUDA-graph), bench with the graphed sharded step and eager.
N=${1:-2}; TAG=${2:-m2}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 10 200 $TR --master-port 29511 tools/sharded_check.py > $O/sharded_check_${TAG}.log 2>&1
echo "sharded_check rc=$?"; grep -E "loss|OK|Error|error" $O/sharded_check_${TAG}.log | tail -12 | cut -c1-250
CLICA_GRAPH_MULTI=1 timeout -k 10 200 $TR --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 > $O/bench_${TAG}_c2_graph.log 2>&1
echo "bench graph rc=$?"; tail -1 $O/bench_${TAG}_c2_graph.log > $O/bench_${TAG}_c2_graph.json
python -c "
import json; d=json.load(open('$O/bench_${TAG}_c2_graph.json')); print('C2 weak N=$N graph: ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['step_mode'])" || tail -5 $O/bench_${TAG}_c2_graph.log | cut -c1-300
CLICA_GRAPH_MULTI=0 timeout -k 10 200 $TR --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 > $O/bench_${TAG}_c2_eager.log 2>&1
echo "bench eager rc=$?"; tail -1 $O/bench_${TAG}_c2_eager.log > $O/bench_${TAG}_c2_eager.json
python -c "
import json; d=json.load(open('$O/bench_${TAG}_c2_eager.json')); print('C2 weak N=$N eager: ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['step_mode'])" || tail -5 $O/bench_${TAG}_c2_eager.log | cut -c1-300
CLICA_GRAPH_MULTI=1 timeout -k 10 200 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 --workload c3 --scaling strong > $O/bench_${TAG}_c3_graph.log 2>&1
echo "bench c3 rc=$?"; tail -1 $O/bench_${TAG}_c3_graph.log > $O/bench_${TAG}_c3_graph.json
python -c "
import json; d=json.load(open('$O/bench_${TAG}_c3_graph.json')); print('C3 strong N=$N graph: ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['step_mode'])" || tail -5 $O/bench_${TAG}_c3_graph.log | cut -c1-300
