python -m pytest tests -q -m gpu > gpurun_out/t52_full.log 2>&1; tail -3 gpurun_out/t52_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke52.log 2>&1; tail -2 gpurun_out/smoke52.log
python bench.py > gpurun_out/bench52_default.log 2>&1; tail -1 gpurun_out/bench52_default.log | cut -c1-300
