#!/bin/bash
# final validation of the round: GPU suite, smoke(), bench (default flags), then the launch list of the final step
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/tests.log 2>&1
tail -3 gpurun_out/tests.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
( timeout 300 python bench.py ) > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-240
( timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches5.csv python profiles/prof_forward.py 2 ) > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
