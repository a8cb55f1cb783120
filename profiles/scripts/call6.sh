#!/bin/bash
# GPU call: launch list of the step (after the neighbour-cache pre kernel) + ncu --set full of the first launches of each hot kernel
# (one small report per kernel: gpurun_out/ is capped at 64 MiB)
mkdir -p gpurun_out
( time timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches4.csv python profiles/prof_forward.py 2 ) > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
for spec in pre_kernel:2 post_kernel:1 attn_tc2_kernel:2 gemm_tc_kernel:3; do
  k=${spec%%:*}; c=${spec##*:}
  ( time timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c $c -f -o gpurun_out/r01d_$k python profiles/prof_forward.py 1 ) > gpurun_out/ncu_$k.log 2>&1
  tail -4 gpurun_out/ncu_$k.log | head -1
done
du -sh gpurun_out; ls -la gpurun_out | grep ncu-rep
