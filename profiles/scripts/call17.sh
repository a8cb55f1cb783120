#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -x -q -k "reduce_ln or model or exact or tensor_core or rng or segmentor or general or simt or single or nuscenes or batch8" ) > gpurun_out/t_model.log 2>&1
tail -3 gpurun_out/t_model.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-260
