#!/bin/bash
# round 2, call 14: (a) does an L1-bypass build (no ld.global.nc, -dlcm=cg) remove the two-stream race?  (b) attention kernel variants
# (112 registers, cheaper spin loop, tc32 at 3 CTAs/SM), vectorised reduce_ln, reduced-precision mode
mkdir -p gpurun_out
( CDSEG_LIB=$PWD/cdsegnet_b200/libcdseg_b200_cg.so timeout 300 python profiles/debug_arena.py ) > gpurun_out/race_cg.log 2>&1
echo "--- L1-bypass build:"; tail -4 gpurun_out/race_cg.log
( timeout 300 python profiles/debug_arena.py ) > gpurun_out/race_default.log 2>&1
echo "--- default build:"; tail -4 gpurun_out/race_default.log
( timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention or reduce_ln" ) > gpurun_out/t_attn14.log 2>&1
tail -3 gpurun_out/t_attn14.log
( timeout 300 python profiles/time_attention_r2.py ) > gpurun_out/time_attn14.log 2>&1
grep -E "tc3 f16   |tc3 tc32|tc2 f16" gpurun_out/time_attn14.log
( timeout 900 python -m pytest tests/test_gpu_full_config.py -m gpu -q ) > gpurun_out/t_full14.log 2>&1
tail -5 gpurun_out/t_full14.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench14_tc32.log 2>&1
tail -1 gpurun_out/bench14_tc32.log | cut -c1-200
( timeout 300 python bench.py --steps 10 --warmup 3 --attention f16 --gemm fp16 --no-cpu ) > gpurun_out/bench14_fp16.log 2>&1
tail -3 gpurun_out/bench14_fp16.log | cut -c1-300
