python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q -m gpu -k "attn or attention or model or native or schedule" > gpurun_out/t51.log 2>&1; tail -2 gpurun_out/t51.log
python profiles/time_attention_r2.py > gpurun_out/time_attention51.txt 2>&1
for i in 1 2; do
  python bench.py --no-cpu --steps 20 > gpurun_out/bench51.log 2>&1
  echo "bench: $(tail -1 gpurun_out/bench51.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["ms_per_step_median"], d["e2e"]["ms_per_step"], d.get("attention_f16",{}).get("ms_per_step"))')"
done
