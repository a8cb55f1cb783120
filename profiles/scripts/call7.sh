#!/bin/bash
# GPU call: wrapper row (criteria kernels, samplers, inference eval / ddim / forward) + full suite
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_wrapper.py -m gpu -x -q ) > gpurun_out/t_wrap.log 2>&1
tail -25 gpurun_out/t_wrap.log
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/tests.log 2>&1
tail -5 gpurun_out/tests.log
