python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_full_config.py -x -q -m gpu > gpurun_out/t34.log 2>&1; tail -2 gpurun_out/t34.log
python profiles/trace_gemm_deep.py > gpurun_out/trace_gemm_deep34.txt 2>&1; cut -c1-150 gpurun_out/trace_gemm_deep34.txt
for P in 1 0 1 0; do
  CDSEG_NO_LDG256=$P python bench.py --no-cpu --steps 20 > gpurun_out/bench34_no256_${P}.log 2>&1
  echo "no256=$P: $(tail -1 gpurun_out/bench34_no256_${P}.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["ms_per_step_median"], d["e2e"]["ms_per_step"], d.get("attention_f16",{}).get("ms_per_step"))')"
done
