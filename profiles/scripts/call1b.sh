#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/tests.log 2>&1
tail -3 gpurun_out/tests.log
( time timeout 300 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench.log 2>&1
tail -2 gpurun_out/bench.log | cut -c1-300
for fam in p0 p1 p2 p3 e; do
  timeout 120 profiles/microbench/aload 120000 $fam > gpurun_out/aload_$fam.log 2>&1
  echo "aload $fam exit $?"
done
du -sh gpurun_out
