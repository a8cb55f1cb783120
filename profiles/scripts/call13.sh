#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "gemm_tc" ) > gpurun_out/t_gemm.log 2>&1
tail -3 gpurun_out/t_gemm.log
( timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_wrapper.py -m gpu -x -q ) > gpurun_out/t_model.log 2>&1
tail -3 gpurun_out/t_model.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-260
( CDSEG_GEMM_NARROW=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_nonarrow.log 2>&1
tail -1 gpurun_out/bench_nonarrow.log | cut -c1-260
