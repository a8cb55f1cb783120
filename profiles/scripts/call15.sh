#!/bin/bash
mkdir -p gpurun_out
for P in 2 3; do
( CDSEG_ATTN_POLY=$P timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "attention_tcgen05_vs" ) > gpurun_out/t_attn_poly$P.log 2>&1
tail -3 gpurun_out/t_attn_poly$P.log
done
( timeout 300 python profiles/time_attention.py ) > gpurun_out/time_attention.log 2>&1
grep "poly=" gpurun_out/time_attention.log
