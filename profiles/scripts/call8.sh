#!/bin/bash
# GPU call: fragment pipeline tests + full suite + bench with the corrected per-kernel roofline objects
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fragments.py -m gpu -x -q ) > gpurun_out/t_frag.log 2>&1
tail -25 gpurun_out/t_frag.log
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/tests.log 2>&1
tail -5 gpurun_out/tests.log
( timeout 400 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-300
