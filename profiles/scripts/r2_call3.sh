#!/bin/bash
# round 2, call 3: native plan + native feature-phase executor
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "plan or patch" ) > gpurun_out/t_plan.log 2>&1
tail -5 gpurun_out/t_plan.log
( timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q ) > gpurun_out/t_model.log 2>&1
tail -8 gpurun_out/t_model.log
( timeout 900 python -m pytest tests/test_gpu_full_config.py -m gpu -q ) > gpurun_out/t_full.log 2>&1
tail -5 gpurun_out/t_full.log
( timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_full_config.py --deselect tests/test_gpu_model.py ) > gpurun_out/t_rest.log 2>&1
tail -5 gpurun_out/t_rest.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench3.log 2>&1
tail -1 gpurun_out/bench3.log | cut -c1-300
( timeout 200 python profiles/host_overhead_r2.py ) > gpurun_out/host3.log 2>&1
head -4 gpurun_out/host3.log
