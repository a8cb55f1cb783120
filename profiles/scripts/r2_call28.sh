python -m pytest tests/test_gpu_model.py tests/test_gpu_full_config.py -x -q -m gpu > gpurun_out/t28.log 2>&1; tail -3 gpurun_out/t28.log
python bench.py --no-cpu --steps 20 > gpurun_out/bench28.log 2>&1
echo "bench: $(tail -1 gpurun_out/bench28.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["ms_per_step_median"], d["e2e"]["ms_per_step"], d.get("attention_f16",{}).get("ms_per_step"))')"
python bench.py --no-cpu --steps 20 --gemm fp16 --attention f16 > gpurun_out/bench28_fp16.log 2>&1
echo "fp16: $(tail -1 gpurun_out/bench28_fp16.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["ms_per_step_median"], d["e2e"]["ms_per_step"])')"
python profiles/timeline_r2.py tc32 gpurun_out/timeline28.csv > gpurun_out/timeline28.txt 2>&1; grep -v "^     gap\|gap histogram" gpurun_out/timeline28.txt | head -12
