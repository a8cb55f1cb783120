#!/bin/bash
mkdir -p gpurun_out
( time timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches3.csv python profiles/prof_forward.py 2 ) > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log
du -sh gpurun_out
