python -m pytest tests/test_gpu_model.py tests/test_gpu_full_config.py tests/test_gpu_wrapper.py -x -q -m gpu > gpurun_out/t32.log 2>&1; tail -2 gpurun_out/t32.log
for i in 1 2; do
python bench.py --no-cpu --steps 20 > gpurun_out/bench32.log 2>&1
echo "bench: $(tail -1 gpurun_out/bench32.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["ms_per_step_median"], d["e2e"]["ms_per_step"], d.get("attention_f16",{}).get("ms_per_step"))')"
done
python profiles/timeline_r2.py tc32 gpurun_out/timeline32.csv > gpurun_out/timeline32.txt 2>&1; grep -v "^     gap" gpurun_out/timeline32.txt | head -14
python profiles/host_overhead_r2.py tc32 > gpurun_out/host32.txt 2>&1; grep -A6 "host timeline" gpurun_out/host32.txt
