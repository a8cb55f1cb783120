python -m pytest tests/test_gpu_model.py tests/test_gpu_full_config.py tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/t30.log 2>&1; tail -3 gpurun_out/t30.log
for P in 0 1; do
  CDSEG_PDL=$P python bench.py --no-cpu --steps 20 > gpurun_out/bench30_pdl${P}.log 2>&1
  echo "pdl=$P: $(tail -1 gpurun_out/bench30_pdl${P}.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["ms_per_step_median"], d["e2e"]["ms_per_step"], d.get("attention_f16",{}).get("ms_per_step"))')"
done
python profiles/timeline_r2.py tc32 gpurun_out/timeline30.csv > gpurun_out/timeline30.txt 2>&1; grep -v "^     gap" gpurun_out/timeline30.txt | head -14
