python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_full_config.py -x -q -m gpu > gpurun_out/t35.log 2>&1; tail -2 gpurun_out/t35.log
python profiles/trace_gemm_deep.py > gpurun_out/trace_gemm_deep35.txt 2>&1; cut -c1-120 gpurun_out/trace_gemm_deep35.txt
for P in 1 0 1 0; do
  CDSEG_NO_FAST_EPI=$P python bench.py --no-cpu --steps 20 > gpurun_out/bench35_nofast_${P}.log 2>&1
  echo "nofast=$P: $(tail -1 gpurun_out/bench35_nofast_${P}.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["ms_per_step_median"], d["e2e"]["ms_per_step"], d.get("attention_f16",{}).get("ms_per_step"))')"
done
