#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_wrapper.py -m gpu -x -q ) > gpurun_out/t_wrap.log 2>&1
tail -25 gpurun_out/t_wrap.log
