#!/bin/bash
# round 2, call 13: full GPU suite on the single-stream native executor, bench in three precision modes + other workloads, launch list, ncu of the HBM-side kernels
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/tests13.log 2>&1
tail -4 gpurun_out/tests13.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke13.log 2>&1
tail -3 gpurun_out/smoke13.log
( timeout 300 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench13_tc32.log 2>&1
tail -1 gpurun_out/bench13_tc32.log | cut -c1-200
( timeout 300 python bench.py --steps 10 --warmup 3 --attention f16 --gemm fp16 --no-cpu ) > gpurun_out/bench13_fp16.log 2>&1
tail -1 gpurun_out/bench13_fp16.log | cut -c1-200
for w in nuscenes scannet200 batch8; do
( timeout 400 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu ) > gpurun_out/bench13_$w.log 2>&1
tail -1 gpurun_out/bench13_$w.log | cut -c1-200
done
( timeout 200 python profiles/host_overhead_r2.py ) > gpurun_out/host13.log 2>&1
head -4 gpurun_out/host13.log
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2b.csv python profiles/prof_forward.py 2 ) > gpurun_out/ncu_list13.log 2>&1
tail -1 gpurun_out/ncu_list13.log
( timeout 600 ncu --set full --clock-control none -k regex:"encode_kernel|rs_hist|rs_scan|rs_scatter|pool_flag|pool_blkscan|pool_write|pool_reduce|gather_rows|unpool_add|pack_heads|pack_split|patch_maps|renumber|nbr_lookup|hash_insert" -o gpurun_out/r02_hbm_kernels -f python profiles/ncu_hbm_kernels.py ) > gpurun_out/ncu_hbm.log 2>&1
tail -2 gpurun_out/ncu_hbm.log
