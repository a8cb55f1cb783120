#!/bin/bash
# round-end style validation: GPU suite, smoke(), bench (both arms, short reference arm)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/tests.log 2>&1
tail -4 gpurun_out/tests.log
( time timeout 600 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
tail -4 gpurun_out/smoke.log
( time timeout 600 python bench.py ) > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-240
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1
tail -4 gpurun_out/bench_ref.log | cut -c1-400
