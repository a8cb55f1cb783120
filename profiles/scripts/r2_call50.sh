for L in "" hint "" hint; do
  if [ -z "$L" ]; then unset CDSEG_LIB; else export CDSEG_LIB=$PWD/cdsegnet_b200/libcdseg_b200_hint.so; fi
  python bench.py --no-cpu --steps 20 > gpurun_out/bench50_$L.log 2>&1
  echo "lib=${L:-default}: $(tail -1 gpurun_out/bench50_$L.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["ms_per_step_median"], d["e2e"]["ms_per_step"], d.get("attention_f16",{}).get("ms_per_step"), d["parity"])')"
done
