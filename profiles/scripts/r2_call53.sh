python bench.py --no-cpu --attention f16 --gemm fp16 > gpurun_out/bench53_fp16.log 2>&1; tail -1 gpurun_out/bench53_fp16.log | cut -c1-200
for W in scannet200 nuscenes batch8; do python bench.py --no-cpu --workload $W --steps 5 > gpurun_out/bench53_$W.log 2>&1; tail -1 gpurun_out/bench53_$W.log | cut -c1-200; done
