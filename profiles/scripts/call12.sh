#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench2.log 2>&1
tail -1 gpurun_out/bench2.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N=2', d['value'], d['ms_per_step'], d.get('ms_per_step_by_rank'), d['clocks'], d['e2e']['ms_per_step'])"
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench1.log 2>&1
tail -1 gpurun_out/bench1.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N=1', d['value'], d['ms_per_step'], d['clocks'], d['e2e']['ms_per_step'])"
nproc; python profiles/host_overhead.py 2>&1 | head -3
