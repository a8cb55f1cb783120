#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -x -q -k "reduce_ln or exact_mode" ) > gpurun_out/t_rl.log 2>&1
tail -2 gpurun_out/t_rl.log
( timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_rl.log 2>&1
tail -1 gpurun_out/bench_rl.log | cut -c1-230
