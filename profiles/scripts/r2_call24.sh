#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/tests24.log 2>&1
tail -4 gpurun_out/tests24.log
( timeout 400 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench24_tc32.log 2>&1
tail -1 gpurun_out/bench24_tc32.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:round(d[k],3) for k in ('ms_per_step','ms_per_step_median','ms_per_step_min','ms_per_step_max')}, 'e2e', round(d['e2e']['ms_per_step'],3), 'f16', round(d['attention_f16']['ms_per_step'],3), 'parity', d['parity']['max_abs'])"
( timeout 300 python bench.py --steps 10 --warmup 3 --attention f16 --gemm fp16 --no-cpu ) > gpurun_out/bench24_fp16.log 2>&1
tail -1 gpurun_out/bench24_fp16.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fp16:', {k:round(d[k],3) for k in ('ms_per_step','ms_per_step_median')}, 'e2e', round(d['e2e']['ms_per_step'],3))"
