python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_full_config.py -x -q -m gpu > gpurun_out/t36.log 2>&1; tail -2 gpurun_out/t36.log
python profiles/trace_gemm_deep.py > gpurun_out/trace_gemm_deep36.txt 2>&1; cut -c1-120 gpurun_out/trace_gemm_deep36.txt
for i in 1 2; do
  python bench.py --no-cpu --steps 20 > gpurun_out/bench36.log 2>&1
  echo "bench: $(tail -1 gpurun_out/bench36.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["ms_per_step_median"], d["e2e"]["ms_per_step"], d.get("attention_f16",{}).get("ms_per_step"))')"
done
