for S in 0 1; do
  CDSEG_SIDE_CTAS=$S python bench.py --no-cpu --steps 20 > gpurun_out/bench29_s${S}.log 2>&1
  echo "side_ctas=$S: $(tail -1 gpurun_out/bench29_s${S}.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["ms_per_step_median"], d["e2e"]["ms_per_step"], d.get("attention_f16",{}).get("ms_per_step"), d["parity"] if "parity" in d else "")')"
done
python profiles/timeline_r2.py tc32 gpurun_out/timeline29.csv > gpurun_out/timeline29.txt 2>&1; grep -v "^     gap\|gap histogram" gpurun_out/timeline29.txt | head -12
