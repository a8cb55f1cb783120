#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "skinny or gemm_tc_linear" ) > gpurun_out/t19a.log 2>&1
tail -3 gpurun_out/t19a.log
( timeout 900 python -m pytest tests/test_gpu_full_config.py tests/test_gpu_model.py tests/test_gpu_wrapper.py -m gpu -q ) > gpurun_out/t19b.log 2>&1
tail -3 gpurun_out/t19b.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench19_tc32.log 2>&1
tail -1 gpurun_out/bench19_tc32.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:round(d[k],3) for k in ('ms_per_step','ms_per_step_median','ms_per_step_min','ms_per_step_max')}, 'e2e', round(d['e2e']['ms_per_step'],3), 'f16', round(d['attention_f16']['ms_per_step'],3))"
( CDSEG_NO_SKINNY=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench19_noskinny.log 2>&1
tail -1 gpurun_out/bench19_noskinny.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no skinny:', {k:round(d[k],3) for k in ('ms_per_step','ms_per_step_median')})"
