python -m pytest tests -m gpu -q 2>&1 | tail -6
python tools/prof_step.py 5 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 7 -c 28 --csv --log-file gpurun_out/launches_r02.csv python tools/prof_step.py 5 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log; wc -l gpurun_out/launches_r02.csv
