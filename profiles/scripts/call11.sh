#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "pre_attn" ) > gpurun_out/t_pre.log 2>&1
tail -2 gpurun_out/t_pre.log
( timeout 200 python profiles/time_fused.py ) > gpurun_out/time_fused.log 2>&1
grep pre_attn gpurun_out/time_fused.log
( timeout 200 python profiles/trace_pre.py ) > gpurun_out/trace_pre.log 2>&1
grep -A2 "CTA 0" gpurun_out/trace_pre.log | cut -c1-330
