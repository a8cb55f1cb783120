python -m pytest tests/test_gpu_model.py tests/test_gpu_full_config.py tests/test_gpu_wrapper.py -x -q -m gpu > gpurun_out/t27.log 2>&1; tail -3 gpurun_out/t27.log
for P in 0 1; do for A in 0 1; do
  CDSEG_PRIORITY_STREAMS=$P CDSEG_PLAN_AUX=$A python bench.py --no-cpu --steps 20 > gpurun_out/bench27_p${P}a${A}.log 2>&1
  echo "prio=$P aux=$A: $(tail -1 gpurun_out/bench27_p${P}a${A}.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["ms_per_step_median"], d["e2e"]["ms_per_step"], d.get("attention_f16",{}).get("ms_per_step"))')"
done; done
python profiles/timeline_r2.py tc32 gpurun_out/timeline27.csv > gpurun_out/timeline27.txt 2>&1; grep -v "^     gap\|gap histogram" gpurun_out/timeline27.txt | head -30
python profiles/host_overhead_r2.py tc32 > gpurun_out/host27.txt 2>&1; grep -A8 "host timeline\|overlap_streams=\|host per forward" gpurun_out/host27.txt | head -40
