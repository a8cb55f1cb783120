set -x
python -m pytest tests -q -m gpu > gpurun_out/t43_full.log 2>&1; tail -3 gpurun_out/t43_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke43.log 2>&1; tail -2 gpurun_out/smoke43.log
python bench.py > gpurun_out/bench43_default.log 2>&1; tail -1 gpurun_out/bench43_default.log | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench43_reference.log 2>&1; tail -1 gpurun_out/bench43_reference.log | cut -c1-400
python bench.py --no-cpu --attention f16 > gpurun_out/bench43_f16.log 2>&1; tail -1 gpurun_out/bench43_f16.log | cut -c1-300
python bench.py --no-cpu --attention f16 --gemm fp16 > gpurun_out/bench43_fp16.log 2>&1; tail -1 gpurun_out/bench43_fp16.log | cut -c1-300
for W in scannet200 nuscenes batch8; do python bench.py --no-cpu --workload $W --steps 5 > gpurun_out/bench43_$W.log 2>&1; tail -1 gpurun_out/bench43_$W.log | cut -c1-300; done
