timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc3 -s 4 -c 1 -f -o gpurun_out/r02b_attn_tc32 python profiles/ncu_attn_r2.py > gpurun_out/ncu49a.log 2>&1; tail -2 gpurun_out/ncu49a.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pre_kernel -s 1 -c 1 -f -o gpurun_out/r02b_pre python profiles/time_fused.py > gpurun_out/ncu49b.log 2>&1; tail -2 gpurun_out/ncu49b.log
timeout 300 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02b_gemm_deep python profiles/trace_gemm_deep.py > gpurun_out/ncu49c.log 2>&1; tail -2 gpurun_out/ncu49c.log
ls -la gpurun_out/*.ncu-rep
python -m pytest tests -q -m gpu > gpurun_out/t49_full.log 2>&1; tail -3 gpurun_out/t49_full.log
python bench.py > gpurun_out/bench49_default.log 2>&1; tail -1 gpurun_out/bench49_default.log | cut -c1-300
