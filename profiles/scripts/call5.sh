#!/bin/bash
# GPU call: neighbour-cache version of the fused pre-attention kernel -- plan + kernel parity, timing, full suite, bench
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "pre_attn or conv_tile_plan" ) > gpurun_out/t_pre.log 2>&1
tail -15 gpurun_out/t_pre.log
( timeout 200 python profiles/time_fused.py ) > gpurun_out/time_fused.log 2>&1
tail -8 gpurun_out/time_fused.log
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/tests.log 2>&1
tail -3 gpurun_out/tests.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-400
