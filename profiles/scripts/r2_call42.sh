python profiles/timeline_r2.py tc32 gpurun_out/timeline42.csv > gpurun_out/timeline42.txt 2>&1; grep -v "^     gap" gpurun_out/timeline42.txt | head -24
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches42.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu42.log 2>&1
python profiles/launch_summary.py gpurun_out/launches42.csv 45 > gpurun_out/launches42.md 2>&1; head -36 gpurun_out/launches42.md
