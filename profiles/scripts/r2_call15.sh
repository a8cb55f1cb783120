#!/bin/bash
mkdir -p gpurun_out
( CDSEG_LIB=$PWD/cdsegnet_b200/libcdseg_b200_cg.so timeout 300 python profiles/debug_arena.py ) > gpurun_out/race_cg.log 2>&1
echo "--- L1-bypass build:"; tail -4 gpurun_out/race_cg.log
( timeout 300 python profiles/debug_arena.py ) > gpurun_out/race_default.log 2>&1
echo "--- default build:"; tail -4 gpurun_out/race_default.log
for s in 0 20 64 200; do CDSEG_ATTN_SLEEP=$s timeout 100 python profiles/time_attention_ab.py; done 2>&1 | tee gpurun_out/attn_ab.log
for v in LB TC32_DEFER; do for s in 0 64; do CDSEG_LIB=$PWD/cdsegnet_b200/libcdseg_b200_$v.so CDSEG_ATTN_SLEEP=$s timeout 100 python profiles/time_attention_ab.py; done; done 2>&1 | tee -a gpurun_out/attn_ab.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench15_tc32.log 2>&1
tail -1 gpurun_out/bench15_tc32.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ms_per_step','ms_per_step_median','ms_per_step_min','ms_per_step_max')}, d['e2e']['ms_per_step'])"
