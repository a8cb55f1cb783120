#!/bin/bash
# GPU call: pre_kernel variants -- parity, timing, model tests, bench
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "pre_attn" ) > gpurun_out/t_pre.log 2>&1
tail -4 gpurun_out/t_pre.log
( timeout 200 python profiles/time_fused.py ) > gpurun_out/time_fused.log 2>&1
grep pre_attn gpurun_out/time_fused.log
( timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q ) > gpurun_out/t_model.log 2>&1
tail -3 gpurun_out/t_model.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-260
