#!/bin/bash
# GPU call: sanity (tests + bench), A-load / epilogue microbenchmarks, ncu --set full over the first forward's own kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/tests.log 2>&1
tail -3 gpurun_out/tests.log
( time timeout 300 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench.log 2>&1
tail -2 gpurun_out/bench.log | cut -c1-400
for fam in p0 p1 p2 p3 e; do
  timeout 120 profiles/microbench/aload 120000 $fam > gpurun_out/aload_$fam.log 2>&1
  echo "aload $fam exit $?"
done
K='regex:encode_kernel|rs_hist_kernel|rs_scan_kernel|rs_scatter_kernel|patch_maps_kernel|pool_flag|pool_blkscan|pool_write|pool_reduce|hash_insert|nbr_lookup|gemm_tc_kernel|add_layernorm|small_linear|pack_heads|attn_tc2|unpool_add'
( time timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -c 140 -o gpurun_out/r01b_full python profiles/prof_forward.py 1 ) > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
