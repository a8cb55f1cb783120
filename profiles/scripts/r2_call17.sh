#!/bin/bash
# round 2, call 17: consolidation -- full GPU suite, smoke, bench (tc32 headline + cpu baseline + parity), launch list, ncu captures
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/tests17.log 2>&1
tail -4 gpurun_out/tests17.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke17.log 2>&1
tail -3 gpurun_out/smoke17.log
( timeout 400 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench17_tc32.log 2>&1
tail -1 gpurun_out/bench17_tc32.log | cut -c1-200
( timeout 300 python bench.py --steps 10 --warmup 3 --attention f16 --gemm fp16 --no-cpu ) > gpurun_out/bench17_fp16.log 2>&1
tail -1 gpurun_out/bench17_fp16.log | cut -c1-200
( timeout 300 python bench.py --steps 10 --warmup 3 --attention f16 --no-cpu ) > gpurun_out/bench17_f16.log 2>&1
tail -1 gpurun_out/bench17_f16.log | cut -c1-200
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2c.csv python profiles/prof_forward.py 2 ) > gpurun_out/ncu_list17.log 2>&1
tail -1 gpurun_out/ncu_list17.log
# one launch each of the hot kernels inside the real forward (stage 0): attention (tc32), pre, post ; f16 attention from the op script
( timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_tc3 -s 37 -c 1 -o gpurun_out/r02_attn_tc32 -f python profiles/prof_forward.py 2 tc32 ) > gpurun_out/ncu_a.log 2>&1
( timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_tc3 -s 37 -c 1 -o gpurun_out/r02_attn_f16 -f python profiles/prof_forward.py 2 f16 ) >> gpurun_out/ncu_a.log 2>&1
( timeout 300 ncu --set full --import-source on --clock-control none -k regex:pre_kernel -s 22 -c 1 -o gpurun_out/r02_pre -f python profiles/prof_forward.py 2 tc32 ) >> gpurun_out/ncu_a.log 2>&1
( timeout 300 ncu --set full --clock-control none -k regex:post_kernel -s 22 -c 1 -o gpurun_out/r02_post -f python profiles/prof_forward.py 2 tc32 ) >> gpurun_out/ncu_a.log 2>&1
grep -c "Report" gpurun_out/ncu_a.log
# HBM-side kernels: one capture per kernel name (-c limits the report size)
( timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --clock-control none -k regex:"encode_kernel|rs_hist|rs_scan|rs_scatter|pool_flag|pool_blkscan|pool_write|pool_reduce|gather_rows|unpool_add|pack_heads|pack_split|patch_maps|renumber|nbr_lookup|hash_insert" -c 80 -o gpurun_out/r02_hbm_kernels -f python profiles/ncu_hbm_kernels.py ) > gpurun_out/ncu_hbm.log 2>&1
tail -2 gpurun_out/ncu_hbm.log
ls -la gpurun_out/*.ncu-rep
du -sh gpurun_out
