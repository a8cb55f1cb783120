#!/bin/bash
# round 2, call 1: new attention kernel (tc3: f16 + tc32 modes), full-config parity tests, bench with parity, host overhead, launch list
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" ) > gpurun_out/t_attn.log 2>&1
tail -3 gpurun_out/t_attn.log
( timeout 300 python profiles/time_attention_r2.py ) > gpurun_out/time_attn.log 2>&1
tail -40 gpurun_out/time_attn.log
( timeout 900 python -m pytest tests/test_gpu_full_config.py tests/test_gpu_model.py -m gpu -q ) > gpurun_out/t_full.log 2>&1
tail -15 gpurun_out/t_full.log
( timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_full_config.py --deselect tests/test_gpu_model.py -k "not attention" ) > gpurun_out/t_rest.log 2>&1
tail -3 gpurun_out/t_rest.log
( timeout 300 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench1.log 2>&1
tail -1 gpurun_out/bench1.log | cut -c1-600
( timeout 200 python profiles/host_overhead_r2.py ) > gpurun_out/host.log 2>&1
head -8 gpurun_out/host.log
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2a.csv python profiles/prof_forward.py 2 ) > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
