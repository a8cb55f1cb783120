#!/bin/bash
# GPU call: knn_query vs oracle + the reference's own kernel (oracle/_ref/libref_knn.so), full suite
mkdir -p gpurun_out
ls -la oracle/_ref
( timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -x -q ) > gpurun_out/t_knn.log 2>&1
tail -25 gpurun_out/t_knn.log
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/tests.log 2>&1
tail -5 gpurun_out/tests.log
