python -m pytest tests/test_gpu_model.py tests/test_gpu_full_config.py -x -q -m gpu > gpurun_out/t31.log 2>&1; tail -2 gpurun_out/t31.log
for P in 0 1 0 1; do
  CDSEG_PDL=$P python bench.py --no-cpu --steps 20 > gpurun_out/bench31_pdl${P}.log 2>&1
  echo "pdl=$P: $(tail -1 gpurun_out/bench31_pdl${P}.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["ms_per_step_median"], d["e2e"]["ms_per_step"], d.get("attention_f16",{}).get("ms_per_step"))')"
done
