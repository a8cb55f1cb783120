#!/bin/bash
# round 2, call 2: what bounds the attention kernel?  what-if variants + ncu --set full with source counters
mkdir -p gpurun_out
( timeout 200 python profiles/attn_whatif.py ) > gpurun_out/whatif.log 2>&1
cat gpurun_out/whatif.log
( timeout 400 ncu --set full --import-source on --clock-control none -k regex:attn_tc3 -s 2 -c 1 -o gpurun_out/r02_attn_tc3_f16 -f python profiles/ncu_attn_r2.py ) > gpurun_out/ncu_a.log 2>&1
tail -2 gpurun_out/ncu_a.log
( timeout 400 ncu --set full --import-source on --clock-control none -k regex:attn_tc3 -s 5 -c 1 -o gpurun_out/r02_attn_tc3_tc32 -f python profiles/ncu_attn_r2.py ) > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
ls -la gpurun_out/*.ncu-rep
